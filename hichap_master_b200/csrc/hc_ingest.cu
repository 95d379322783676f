// Valid-pair text ingest (host code, multithreaded): the step right before the binning kernel.
// Replaces the interpreted per-line parsing of HiCHap/matrixBuilding.py:573-580 (23-column
// *_Valid.bed written by filtering.py:398; columns documented at filtering.py:16-47) and
// :822-829 / :1131-1141 (4/5-column allelic beds written by filtering.py:913, :1127-1234), including
// `lstrip('chr')` (a character-SET strip), the chromosome filter (:577) and the concatenation of
// several files (`cat`, :307-313).  Files are mmap'ed, cut into newline-aligned chunks, parsed by a
// pool of std::threads into per-chunk columns and concatenated in file order.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "hc_common.cuh"

namespace {

struct Mapped { const char* p = nullptr; size_t n = 0; int fd = -1; };

bool map_file(const char* path, Mapped* m) {
    m->fd = open(path, O_RDONLY);
    if (m->fd < 0) return false;
    struct stat st;
    if (fstat(m->fd, &st) != 0) { close(m->fd); return false; }
    m->n = (size_t)st.st_size;
    if (m->n == 0) { m->p = nullptr; return true; }
    void* a = mmap(nullptr, m->n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, m->fd, 0);
    if (a == MAP_FAILED) { close(m->fd); return false; }
    madvise(a, m->n, MADV_SEQUENTIAL);
    m->p = (const char*)a;
    return true;
}
void unmap_file(Mapped* m) {
    if (m->p) munmap((void*)m->p, m->n);
    if (m->fd >= 0) close(m->fd);
}

// whitespace as str.split() sees it inside a line ('\n' never occurs inside one)
struct WsTable {
    bool t[256];
    WsTable() { for (int i = 0; i < 256; ++i) t[i] = i == ' ' || i == '\t' || i == '\r' || i == '\f' || i == '\v'; }
};
const WsTable g_ws;
inline bool is_ws(char c) { return g_ws.t[(unsigned char)c]; }

struct ChromTable {
    std::unordered_map<std::string, int> id;   // stripped name -> index into the sorted chromosome table

    std::vector<std::string> explicit_names;   // the `chroms` list without '#'
    bool keep_all = false, keep_numeric = false;
    // matrixBuilding.py:360: (not chroms) or (c.isdigit() and '#' in chroms) or (c in chroms)
    bool passes(const std::string& c) const {
        if (keep_all) return true;
        if (keep_numeric && !c.empty()) {
            bool dig = true;
            for (char ch : c) if (ch < '0' || ch > '9') { dig = false; break; }
            if (dig) return true;
        }
        for (const auto& e : explicit_names) if (e == c) return true;
        return false;
    }
};

struct ChunkOut {
    std::vector<int32_t> c1, p1, c2, p2;
    std::vector<uint8_t> mark;
    std::string error;    // first error in this chunk
    int error_kind = 0;   // 1 = KeyError (chromosome not in genome), 2 = ValueError / malformed line
};

// names of <= 7 bytes packed into one word (length in the top byte): a per-chunk cache of the few
// distinct chromosome names avoids a std::string + hash lookup per field
inline uint64_t pack8(const char* s, size_t n) {
    uint64_t w = 0;
    memcpy(&w, s, n);
    return w | ((uint64_t)n << 56);
}

// returns -1 for "dropped by the filter", -2 for KeyError
int chrom_id(const ChromTable& T, std::vector<std::pair<uint64_t, int>>& cache, const char* s, size_t n,
             std::string* keyerr) {
    size_t i = 0;
    while (i < n && (s[i] == 'c' || s[i] == 'h' || s[i] == 'r')) ++i;   // str.lstrip('chr')
    const size_t m = n - i;
    uint64_t key = 0;
    if (m <= 7) {                                                        // 7 bytes + length tag in the top byte
        key = pack8(s + i, m);
        for (const auto& e : cache) if (e.first == key) return e.second;
    }
    std::string c(s + i, m);
    int r;
    if (!T.passes(c)) r = -1;
    else {
        auto it = T.id.find(c);
        if (it == T.id.end()) { *keyerr = c; return -2; }
        r = it->second;
    }
    if (m <= 7 && cache.size() < 4096) cache.emplace_back(key, r);
    return r;
}

bool parse_int(const char* s, size_t n, long long* out) {
    size_t i = 0;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
    if (i >= n) return false;
    long long v = 0;
    for (; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + (s[i] - '0');
        if (v > (1ll << 40)) return false;
    }
    *out = neg ? -v : v;
    return true;
}

void parse_chunk(const char* b, const char* e, int layout, const ChromTable& T, ChunkOut* out) {
    const char* fs[24]; size_t fl[24];
    std::vector<std::pair<uint64_t, int>> cache;
    const size_t guess = (size_t)(e - b) / (layout == 0 ? 60 : 24) + 16;
    out->c1.reserve(guess); out->p1.reserve(guess); out->c2.reserve(guess); out->p2.reserve(guess);
    if (layout != 0) out->mark.reserve(guess);
    while (b < e) {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;
        // split on runs of whitespace: only the first `need` fields are located one by one (the valid-pair layout
        // uses fields 1, 6, 8, 13 of 23); the last field (the allelic mark) is found from the end of the line
        const int need = layout == 0 ? 14 : 4;
        int nf = 0;
        const char* last_s = nullptr; size_t last_n = 0;
        const char* p = b;
        while (p < le && nf < need) {
            while (p < le && is_ws(*p)) ++p;
            if (p >= le) break;
            const char* q = p;
            while (q < le && !is_ws(*q)) ++q;
            fs[nf] = p; fl[nf] = (size_t)(q - p);
            last_s = p; last_n = (size_t)(q - p);
            ++nf;
            p = q;
        }
        if (layout != 0 && nf == need) {          // line[-1]: the last whitespace-separated token of the line
            const char* r = le;
            while (r > p && is_ws(r[-1])) --r;
            if (r > p) {                          // there is something after field 3
                const char* l = r;
                while (l > p && !is_ws(l[-1])) --l;
                last_s = l; last_n = (size_t)(r - l);
            }
        }
        b = nl ? nl + 1 : e;
        if (nf == 0) continue;    // blank line
        int ia, ipa, ib, ipb;
        if (layout == 0) { ia = 1; ipa = 6; ib = 8; ipb = 13; } else { ia = 0; ipa = 1; ib = 2; ipb = 3; }
        if (nf <= ipb) {
            if (out->error.empty()) { out->error = "line with too few columns"; out->error_kind = 2; }
            continue;
        }
        std::string ke;
        const int a = chrom_id(T, cache, fs[ia], fl[ia], &ke);
        const int c = chrom_id(T, cache, fs[ib], fl[ib], &ke);
        if (a == -1 || c == -1) continue;                 // dropped by the chromosome filter (:580)
        if (a == -2 || c == -2) {
            if (out->error.empty()) { out->error = ke; out->error_kind = 1; }
            continue;
        }
        long long x, y;
        if (!parse_int(fs[ipa], fl[ipa], &x) || !parse_int(fs[ipb], fl[ipb], &y) || x > 2147483647ll || y > 2147483647ll ||
            x < -2147483648ll || y < -2147483648ll) {
            if (out->error.empty()) { out->error = "invalid integer position"; out->error_kind = 2; }
            continue;
        }
        out->c1.push_back(a); out->p1.push_back((int32_t)x); out->c2.push_back(c); out->p2.push_back((int32_t)y);
        if (layout != 0) {
            uint8_t mk = 3;
            if (last_n == 4 && memcmp(last_s, "Both", 4) == 0) mk = 0;
            else if (last_n == 2 && memcmp(last_s, "R1", 2) == 0) mk = 1;
            else if (last_n == 2 && memcmp(last_s, "R2", 2) == 0) mk = 2;
            out->mark.push_back(mk);
        }
    }
}

struct ParseJob {
    std::vector<Mapped> files;
    std::vector<std::pair<const char*, const char*>> chunks;
    std::vector<ChunkOut> outs;
    int layout = 0;
    int nthreads = 0;
};

thread_local ParseJob* g_job = nullptr;   // result of the last hc_ingest_parse on this thread

}  // namespace

// Parse `npaths` files (in order).  layout: 0 = 23-column valid bed, 1 = 4/5-column allelic bed.
// chrom_names[i] (already lstrip('chr')-stripped) -> index i for the chromosomes of the genome table;
// filter_names: the `chroms` list (may contain "#"; n = 0 keeps everything).  Returns the number of
// kept pairs in *h_npairs; the columns are then fetched with hc_ingest_fetch (same thread).
// Error codes: HC_ERR_ARG with hc_last_error() = "KeyError: <chrom>" / "ValueError: ..." / "IOError: ...".
extern "C" int hc_ingest_parse(const char* const* paths, int32_t npaths, int32_t layout,
                               const char* const* chrom_names, int32_t nchrom,
                               const char* const* filter_names, int32_t nfilter, int32_t nthreads,
                               int64_t* h_npairs) {
    HC_REQUIRE(npaths >= 0 && (layout == 0 || layout == 1) && h_npairs != nullptr, "arguments");
    delete g_job;
    g_job = new ParseJob();
    ParseJob& J = *g_job;
    J.layout = layout;
    ChromTable T;
    for (int i = 0; i < nchrom; ++i) T.id[chrom_names[i]] = i;
    T.keep_all = nfilter == 0;
    for (int i = 0; i < nfilter; ++i) {
        if (strcmp(filter_names[i], "#") == 0) T.keep_numeric = true;
        else T.explicit_names.push_back(filter_names[i]);
    }
    J.files.resize(npaths);
    for (int i = 0; i < npaths; ++i) {
        if (!map_file(paths[i], &J.files[i])) {
            hc_set_error("IOError: cannot open %s", paths[i]);
            for (int k = 0; k < i; ++k) unmap_file(&J.files[k]);
            delete g_job; g_job = nullptr;
            return HC_ERR_ARG;
        }
    }
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    J.nthreads = nthreads;
    const size_t target = 8u << 20;   // ~8 MB newline-aligned chunks, in file order
    for (auto& f : J.files) {
        const char* b = f.p;
        const char* e = f.p + f.n;
        while (b && b < e) {
            const char* c = b + target < e ? b + target : e;
            if (c < e) { const char* nl = (const char*)memchr(c, '\n', (size_t)(e - c)); c = nl ? nl + 1 : e; }
            J.chunks.emplace_back(b, c);
            b = c;
        }
    }
    J.outs.resize(J.chunks.size());
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= J.chunks.size()) break;
            parse_chunk(J.chunks[i].first, J.chunks[i].second, layout, T, &J.outs[i]);
        }
    };
    std::vector<std::thread> pool;
    const int nt = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(J.chunks.size(), 1));
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    for (auto& f : J.files) unmap_file(&f);
    J.files.clear();
    int64_t total = 0;
    for (auto& o : J.outs) {
        if (!o.error.empty()) {
            hc_set_error("%s: %s", o.error_kind == 1 ? "KeyError" : "ValueError", o.error.c_str());
            delete g_job; g_job = nullptr;
            return HC_ERR_ARG;
        }
        total += (int64_t)o.c1.size();
    }
    *h_npairs = total;
    return HC_OK;
}

// Copy the parsed columns into caller buffers (npairs entries each; mark may be NULL) and free them.
extern "C" int hc_ingest_fetch(int32_t* c1, int32_t* p1, int32_t* c2, int32_t* p2, uint8_t* mark) {
    HC_REQUIRE(g_job != nullptr, "no parsed data on this thread");
    ParseJob& J = *g_job;
    std::vector<size_t> off(J.outs.size() + 1, 0);
    for (size_t i = 0; i < J.outs.size(); ++i) off[i + 1] = off[i] + J.outs[i].c1.size();
    const bool want_mark = mark != nullptr && J.layout != 0;
    std::atomic<size_t> next{0};
    auto worker = [&]() {       // chunk-parallel: the destination pages are touched for the first time here
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= J.outs.size()) break;
            ChunkOut& o = J.outs[i];
            const size_t n = o.c1.size();
            if (n) {
                memcpy(c1 + off[i], o.c1.data(), n * 4); memcpy(p1 + off[i], o.p1.data(), n * 4);
                memcpy(c2 + off[i], o.c2.data(), n * 4); memcpy(p2 + off[i], o.p2.data(), n * 4);
                if (want_mark) memcpy(mark + off[i], o.mark.data(), n);
            }
            ChunkOut().c1.swap(o.c1); ChunkOut().p1.swap(o.p1); ChunkOut().c2.swap(o.c2); ChunkOut().p2.swap(o.p2);
        }
    };
    int nt = J.nthreads > 0 ? J.nthreads : (int)std::thread::hardware_concurrency();
    nt = (int)std::min<size_t>((size_t)std::max(nt, 1), std::max<size_t>(J.outs.size(), 1));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    delete g_job;
    g_job = nullptr;
    return HC_OK;
}
