// host_fastx.cuh — FASTA / FASTQ front-end: records -> the parseInput blob + seqPair index that every
// other entry point takes (c++/parseInput.h:9-29), so the same batches can come from the formats sequencers and assemblers
// write instead of the project's own 3-line records.  Host code, one pass per file; included inside dpxalign.cu's extern "C" block.
//
//   FASTA  : '>' header line, then sequence lines up to the next '>' (line breaks, '\r' and blanks inside a sequence are dropped)
//   FASTQ  : '@' header, sequence line(s), '+' line, as many quality characters as bases (skipped)
// One file: records alternate reference, query.  Two files: record k of the first is the reference of pair k, record k of the
// second its query.  The blob is "ref\0qry\0ref\0qry\0..." and both arrays are malloc'ed (dpx_free).
#pragma once

static int fastx_slurp(const char* path, std::vector<char>& buf) {
    FILE* f = fopen(path, "rb");
    if (!f) return DPX_ERR_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return DPX_ERR_IO; }
    const long sz = ftell(f);
    if (sz < 0) { fclose(f); return DPX_ERR_IO; }
    rewind(f);
    buf.resize((size_t)sz);
    const size_t got = sz ? fread(buf.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz ? DPX_OK : DPX_ERR_IO;
}

// Appends every record's bases to `blob` (NUL-terminated) and its (offset, length) to `recs`.  Works a line at a time (memchr + one
// bulk append per line), so multi-GB files parse at memory speed.
static int fastx_records(const std::vector<char>& in, std::vector<char>& blob, std::vector<std::pair<size_t, size_t>>& recs) {
    const size_t n = in.size();
    const char* d = in.data();
    size_t i = 0;
    auto skip_blank = [&]() { while (i < n && (d[i] == '\n' || d[i] == '\r' || d[i] == ' ' || d[i] == '\t')) ++i; };
    auto line_end = [&](size_t from) { const void* p = from < n ? memchr(d + from, '\n', n - from) : nullptr; return p ? (size_t)((const char*)p - d) : n; };
    auto skip_line = [&]() { i = line_end(i); if (i < n) ++i; };
    // appends the line starting at i without blanks / '\r', leaves i on the next line
    auto take_line = [&]() {
        const size_t e = line_end(i);
        size_t a = i, b = e;
        while (b > a && (d[b - 1] == '\r' || d[b - 1] == ' ' || d[b - 1] == '\t')) --b;
        if (memchr(d + a, ' ', b - a) || memchr(d + a, '\t', b - a)) { for (size_t k = a; k < b; ++k) if (d[k] != ' ' && d[k] != '\t' && d[k] != '\r') blob.push_back(d[k]); }
        else blob.insert(blob.end(), d + a, d + b);
        i = e < n ? e + 1 : n;
    };
    blob.reserve(blob.size() + n);
    skip_blank();
    while (i < n) {
        const char kind = d[i];
        if (kind != '>' && kind != '@') return DPX_ERR_FORMAT;
        skip_line();                                                 // header
        const size_t start = blob.size();
        if (kind == '>') {
            while (i < n && d[i] != '>') take_line();                // sequence lines up to the next record
        } else {
            while (i < n && d[i] != '+') take_line();                // sequence lines up to the '+' separator (never a base)
            if (i >= n) return DPX_ERR_FORMAT;
            skip_line();                                             // '+' line
            size_t q = 0;
            const size_t len = blob.size() - start;
            while (i < n && q < len) {                               // quality: as many characters as bases, '@' and '>' included
                const size_t e = line_end(i);
                size_t b = e;
                while (b > i && d[b - 1] == '\r') --b;
                q += b - i;
                i = e < n ? e + 1 : n;
            }
            if (q != len) return DPX_ERR_FORMAT;
        }
        recs.emplace_back(start, blob.size() - start);
        blob.push_back('\0');
        skip_blank();
    }
    return DPX_OK;
}

int dpx_parse_fastx(const char* path_refs, const char* path_queries, dpx_seq_pair** pairs_out, char** seq_out, dpx_input_info* info) {
    if (!path_refs || !pairs_out || !seq_out) return DPX_ERR_INVALID;
    *pairs_out = nullptr; *seq_out = nullptr;
    std::vector<char> file, blob;
    std::vector<std::pair<size_t, size_t>> ra, rb;
    { int s = fastx_slurp(path_refs, file); if (s) return s; }
    { int s = fastx_records(file, blob, ra); if (s) return s; }
    if (path_queries) {
        { int s = fastx_slurp(path_queries, file); if (s) return s; }
        { int s = fastx_records(file, blob, rb); if (s) return s; }
        if (ra.size() != rb.size()) return DPX_ERR_FORMAT;          // every reference needs its query
    } else {
        if (ra.size() % 2 != 0) return DPX_ERR_FORMAT;              // records alternate reference, query
        for (size_t k = 0; k < ra.size(); k += 2) { rb.push_back(ra[k + 1]); ra[k / 2] = ra[k]; }
        ra.resize(rb.size());
    }
    if (blob.size() > 0x7fffffffu) return DPX_ERR_RANGE;            // seqPair offsets are int (parseInput.h:9-14)
    const size_t n = std::min<size_t>(ra.size(), 10000000);          // INPUT_CAP, parseInput.cpp:7,102-105
    dpx_seq_pair* idx = (dpx_seq_pair*)malloc(std::max<size_t>(n, 1) * sizeof(dpx_seq_pair));
    char* seq = (char*)malloc(std::max<size_t>(blob.size(), 1));
    if (!idx || !seq) { free(idx); free(seq); return DPX_ERR_NOMEM; }
    if (!blob.empty()) memcpy(seq, blob.data(), blob.size());
    dpx_input_info in{}; in.minReferenceLength = SIZE_MAX; in.minQueryLength = SIZE_MAX;
    for (size_t k = 0; k < n; ++k) {
        idx[k].referenceIdx = (int32_t)ra[k].first; idx[k].referenceSize = (int32_t)ra[k].second;
        idx[k].queryIdx = (int32_t)rb[k].first; idx[k].querySize = (int32_t)rb[k].second;
        in.avgReferenceLength += (double)ra[k].second; in.avgQueryLength += (double)rb[k].second;
        in.maxReferenceLength = std::max(in.maxReferenceLength, ra[k].second); in.minReferenceLength = std::min(in.minReferenceLength, ra[k].second);
        in.maxQueryLength = std::max(in.maxQueryLength, rb[k].second); in.minQueryLength = std::min(in.minQueryLength, rb[k].second);
        in.numCells += ra[k].second * rb[k].second;
    }
    in.numPairs = n; in.numBytes = blob.size();
    if (n) { in.avgReferenceLength /= (double)n; in.avgQueryLength /= (double)n; }
    *pairs_out = idx; *seq_out = seq;
    if (info) *info = in;
    if (n) dpxhost_pack::register_input(seq, blob.size(), idx, n);     // 2-bit sidecar for the one-call path (host_pack.h)
    return DPX_OK;
}
