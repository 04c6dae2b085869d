// mtg-b200 host input: FASTA/FASTQ reader in the style of kseq, mirroring the behaviour of gatb's BankFasta
// (thirdparty/gatb-core/gatb-core/src/gatb/bank/impl/BankFasta.cpp:485-574): '>' or '@' headers, multi-line sequences,
// FASTQ quality skipped by length; comma separated file lists (README.md:166). Plain or gzip files (zlib reads both, like
// gatb's buffered_file_t over gzFile, BankFasta.cpp:52-60).
// getCommentShort = header up to the first whitespace (gatb/bank/api/Sequence.hpp:88).
#pragma once
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mtg {

struct SeqRecord { std::string name, seq; };

class SeqReader {
    gzFile f_ = nullptr;
    std::vector<char> buf_;
    size_t b_ = 0, e_ = 0;
    bool eof_ = false;
    int last_ = 0;
    uint64_t total_ = 0;
    int getc_() {
        if (b_ >= e_) {
            if (eof_) return -1;
            const int r = gzread(f_, buf_.data(), (unsigned)buf_.size());
            if (r < 0) throw std::runtime_error("read error in a sequence file");
            e_ = (size_t)r;
            total_ += e_;
            b_ = 0;
            if (e_ < buf_.size()) eof_ = true;
            if (e_ == 0) return -1;
        }
        return (unsigned char)buf_[b_++];
    }
    // append up to '\n' to s; returns false at EOF with nothing read
    bool getline_(std::string& s, bool append) {
        if (!append) s.clear();
        bool any = false;
        while (true) {
            if (b_ >= e_) { if (getc_() < 0) break; b_--; }
            size_t i = b_;
            while (i < e_ && buf_[i] != '\n') i++;
            s.append(buf_.data() + b_, i - b_);
            any = true;
            if (i < e_) { b_ = i + 1; break; }
            b_ = e_;
        }
        if (!s.empty() && s.back() == '\r') s.pop_back();
        return any;
    }

public:
    explicit SeqReader(const std::string& path) : buf_(1 << 22) {
        f_ = gzopen(path.c_str(), "rb");
        if (!f_) throw std::runtime_error("Cannot open file " + path);
        gzbuffer(f_, 1u << 20);
    }
    ~SeqReader() { if (f_) gzclose(f_); }
    uint64_t bytes_read() const { return total_; }
    bool next(SeqRecord& r) {
        int c;
        if (last_ == 0) {
            while ((c = getc_()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return false;
            last_ = c;
        }
        std::string header;
        if (!getline_(header, false)) return false;
        size_t sp = 0;
        while (sp < header.size() && !isspace((unsigned char)header[sp])) sp++;
        r.name.assign(header, 0, sp);
        r.seq.clear();
        while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            r.seq.push_back((char)c);
            getline_(r.seq, true);
        }
        if (c == '>' || c == '@') last_ = c;
        if (c == '+') {
            std::string q;
            getline_(q, false);  // rest of the '+' line
            size_t qlen = 0;
            while (qlen < r.seq.size() && getline_(q, false)) qlen += q.size();
            last_ = 0;
        }
        if (c == -1) last_ = 0;
        return true;
    }
};

// "File of files" (README.md:166; gatb's BankAlbum, bank/impl/BankAlbum.cpp:48-94, is tried before BankFasta and accepts a text
// file whose every non-blank line names an existing file, :124-170; a bare file name is relative to the album's directory).
// Returns true and the listed paths when `path` is such a file.
inline bool album_paths(const std::string& path, std::vector<std::string>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const int c0 = fgetc(f);
    if (c0 == EOF || c0 == '>' || c0 == '@' || c0 == 0x1f) { fclose(f); return false; }   // sequence text or gzip
    rewind(f);
    std::string dir = ".";
    const size_t slash = path.rfind('/');
    if (slash != std::string::npos) dir = path.substr(0, slash);
    std::vector<std::string> paths;
    char line[4096];
    bool ok = true;
    while (ok && fgets(line, sizeof line, f)) {
        size_t n = strlen(line);
        while (n && isspace((unsigned char)line[n - 1])) line[--n] = 0;
        if (!n) continue;
        for (size_t i = 0; i < n; i++) if ((unsigned char)line[i] < 32) ok = false;   // binary data is not a list of paths
        std::string u(line);
        if (u.find('/') == std::string::npos) u = dir + "/" + u;
        FILE* g = ok ? fopen(u.c_str(), "rb") : nullptr;
        if (!g) { ok = false; break; }
        fclose(g);
        paths.push_back(u);
        if (paths.size() > 100000) ok = false;
    }
    fclose(f);
    if (!ok || paths.empty()) return false;
    out = paths;
    return true;
}
// The files a comma separated -in / -ref list names, albums expanded (recursively, like Bank::open on each line).
inline void expand_uri(const std::string& uri, std::vector<std::string>& files, int depth = 0) {
    if (depth > 8) throw std::runtime_error("file-of-files nested too deeply: " + uri);
    size_t start = 0;
    while (start <= uri.size()) {
        size_t c = uri.find(',', start);
        std::string path = uri.substr(start, c == std::string::npos ? std::string::npos : c - start);
        if (!path.empty()) {
            std::vector<std::string> listed;
            if (album_paths(path, listed)) for (const std::string& p : listed) expand_uri(p, files, depth + 1);
            else files.push_back(path);
        }
        if (c == std::string::npos) break;
        start = c + 1;
    }
}

inline void for_each_sequence(const std::string& uri, const std::function<void(SeqRecord&)>& fn) {
    std::vector<std::string> files;
    expand_uri(uri, files);
    for (const std::string& path : files) {
        SeqReader rd(path);
        SeqRecord r;
        size_t nrec = 0;
        while (rd.next(r)) { fn(r); nrec++; }
        if (!nrec && rd.bytes_read() > 0) throw std::runtime_error("no FASTA/FASTQ record in " + path);
    }
}

// -bed (src/FindBreakpoints.hpp:462-495): the intervals of chromosome `chrom`, in file order. Lines that are empty or start
// with '#' or '@' are ignored; fields are tab separated; field 0 must equal the sequence's short name; begin and end are read
// like std::stoi (leading blanks, sign, then digits; anything after them is ignored, so "140 SNP T -> C" is 140); a line
// is kept when (end - begin) > k in unsigned 64-bit arithmetic. A non-numeric field is an error, as stoi throws there.
inline long bed_stoi(const std::string& f, const std::string& line) {
    char* endp = nullptr;
    errno = 0;
    const long v = strtol(f.c_str(), &endp, 10);
    if (endp == f.c_str()) throw std::runtime_error("bed: not a number in line: " + line);
    if (errno == ERANGE || v > 2147483647L || v < -2147483648L) throw std::runtime_error("bed: number out of range in line: " + line);
    return v;
}
inline std::vector<std::pair<uint64_t, uint64_t>> bed_intervals(const std::string& bed_text, const std::string& chrom, int k) {
    std::vector<std::pair<uint64_t, uint64_t>> iv;
    for (size_t a = 0; a < bed_text.size();) {
        size_t b = bed_text.find('\n', a);
        if (b == std::string::npos) b = bed_text.size();
        const std::string line(bed_text, a, b - a);
        a = b + 1;
        if (line.empty() || line[0] == '#' || line[0] == '@') continue;
        const size_t t1 = line.find('\t');
        if (line.compare(0, t1 == std::string::npos ? line.size() : t1, chrom) != 0) continue;
        const size_t t2 = t1 == std::string::npos ? t1 : line.find('\t', t1 + 1);
        if (t2 == std::string::npos) throw std::runtime_error("bed: fewer than 3 tab-separated fields in line: " + line);
        const size_t t3 = line.find('\t', t2 + 1);
        const uint64_t lo = (uint64_t)bed_stoi(line.substr(t1 + 1, t2 - t1 - 1), line);
        const uint64_t hi = (uint64_t)bed_stoi(line.substr(t2 + 1, t3 == std::string::npos ? t3 : t3 - t2 - 1), line);
        if (hi - lo > (uint64_t)k) iv.emplace_back(lo, hi);
    }
    return iv;
}
inline std::string read_text_file(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open file " + path);
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

}  // namespace mtg
