// textio.cu -- native edge-list reader and feature writer with the reference's wire formats
// (SURVEY.md section 8f row 2).  Host code only: no kernels, no device memory.
//
// Reference being replaced (paths relative to /root/reference/reveal_graph_embedding/):
//   read_adjacency_matrix   datautil/datarw.py:54-120  `src<sep>dst<sep>weight` rows, '#' comments,
//                           node ids renumbered 0..n-1 in first-seen order (source before target),
//                           optional reciprocal edge after each edge (self loops once)
//   write_features          datautil/datarw.py:123-143 one `node_id<sep>community<sep>int(value)` row
//                           per stored entry, row-major
//   get_file_row_generator  common.py:36-49            line.strip().split(separator)
//
// Both are pure-Python per-line / per-entry loops in the reference; around a one-second GPU
// extraction they are the whole wall time of the `arcte` console script (a YouTube-shaped
// result has 5.2e8 entries).  Here the file is parsed / formatted by all host threads.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct arcte_cuda_edge_list {
    std::vector<int64_t> row, col, node_ids;
    std::vector<double> data;
};

namespace arcte {
namespace {

inline bool is_space(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\n' || ch == '\v' || ch == '\f'; }

struct ParsedChunk {
    std::vector<int64_t> src, dst;
    std::vector<double> w;
    int64_t bad_line = -1;  // chunk-relative index of the first unparseable line
    std::string bad_text;
    int64_t lines = 0;
};

// [b, e) -> int64, Python int() style: optional surrounding blanks and sign, decimal digits only.
bool parse_int(const char *b, const char *e, int64_t *out)
{
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; ++b; }
    if (b == e) return false;
    uint64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        const uint64_t d = (uint64_t)(*b - '0');
        if (v > (UINT64_C(0x7fffffffffffffff) - d) / 10) return false;  // would not fit int64
        v = v * 10 + d;
    }
    *out = neg ? -(int64_t)v : (int64_t)v;
    return true;
}

// [b, e) -> double through strtod (correctly rounded like Python's float()); the whole field
// must be consumed.  Hex floats, which Python rejects, are rejected too.
bool parse_float(const char *b, const char *e, double *out)
{
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b == e || e - b > 120) return false;
    char buf[128];
    memcpy(buf, b, (size_t)(e - b));
    buf[e - b] = 0;
    for (const char *p = buf; *p; ++p)
        if (*p == 'x' || *p == 'X') return false;
    char *end = nullptr;
    errno = 0;
    const double v = strtod(buf, &end);
    if (end != buf + (e - b)) return false;
    *out = v;
    return true;
}

void parse_range(const char *beg, const char *end, const std::string &sep, ParsedChunk *pc)
{
    const char *p = beg;
    const size_t sl = sep.size();
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const char *b = p, *e = le;
        while (b < e && is_space(*b)) ++b;   // line.strip()
        while (e > b && is_space(e[-1])) --e;
        const int64_t line_no = pc->lines++;
        p = nl ? nl + 1 : end;
        if (b == e || *b == '#') continue;    // datarw.py:82-83 (blank lines crash the reference; skipped here)
        // split(separator): the first two separators delimit the three fields that are read
        const char *f[4];
        int nf = 0;
        f[nf++] = b;
        const char *q = b;
        while (nf < 4 && q + sl <= e) {
            if (memcmp(q, sep.data(), sl) == 0) {
                q += sl;
                f[nf++] = q;
            } else {
                ++q;
            }
        }
        int64_t s = 0, t = 0;
        double w = 0.0;
        bool ok = nf >= 3;
        if (ok) {
            const char *e2 = nf >= 4 ? f[3] - sl : e;
            ok = parse_int(f[0], f[1] - sl, &s) && parse_int(f[1], f[2] - sl, &t) && parse_float(f[2], e2, &w);
        }
        if (!ok) {
            if (pc->bad_line < 0) {
                pc->bad_line = line_no;
                pc->bad_text.assign(b, (size_t)std::min<ptrdiff_t>(e - b, 80));
            }
            continue;
        }
        pc->src.push_back(s);
        pc->dst.push_back(t);
        pc->w.push_back(w);
    }
}

// Open-addressing map original id -> dense index in first-seen order (datarw.py:88-93).  Key and
// value share a 16-byte slot (one cache miss per probe); callers prefetch the slot of an id a
// few edges ahead, which hides the DRAM latency of the random probes.
struct IdMap {
    struct Slot { int64_t key, val; };
    std::vector<Slot> slots;
    std::vector<int64_t> order;  // dense index -> original id
    uint64_t mask = 0;
    explicit IdMap(size_t expected)
    {
        size_t cap = 16;
        while (cap < expected * 2 + 16) cap <<= 1;
        slots.assign(cap, Slot{0, -1});
        mask = cap - 1;
    }
    static uint64_t mix(uint64_t x)
    {
        x ^= x >> 33; x *= UINT64_C(0xff51afd7ed558ccd); x ^= x >> 33; x *= UINT64_C(0xc4ceb9fe1a85ec53); x ^= x >> 33;
        return x;
    }
    void prefetch(int64_t id) const { __builtin_prefetch(&slots[mix((uint64_t)id) & mask], 1, 1); }
    int64_t get(int64_t id)
    {
        uint64_t h = mix((uint64_t)id) & mask;
        for (;;) {
            Slot &sl = slots[h];
            if (sl.val < 0) {
                sl.key = id;
                sl.val = (int64_t)order.size();
                order.push_back(id);
                return sl.val;
            }
            if (sl.key == id) return sl.val;
            h = (h + 1) & mask;
        }
    }
};

inline int n_digits(uint64_t v)
{
    int d = 1;
    while (v >= 10000) { v /= 10000; d += 4; }
    if (v >= 1000) return d + 3;
    if (v >= 100) return d + 2;
    if (v >= 10) return d + 1;
    return d;
}
inline int int_len(int64_t v) { return v < 0 ? 1 + n_digits((uint64_t)(-(v + 1)) + 1) : n_digits((uint64_t)v); }
inline char *put_int(char *p, int64_t v)
{
    uint64_t u = (uint64_t)v;
    if (v < 0) { *p++ = '-'; u = (uint64_t)(-(v + 1)) + 1; }
    const int d = n_digits(u);
    for (int i = d - 1; i >= 0; --i) { p[i] = (char)('0' + u % 10); u /= 10; }
    return p + d;
}

int pick_threads(int n_threads)
{
    if (n_threads > 0) return n_threads;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

}  // namespace
}  // namespace arcte

using namespace arcte;

extern "C" {

int arcte_cuda_io_read_edge_list(const char *path, const char *separator, int undirected, int n_threads,
                                 arcte_cuda_edge_list **out, int64_t *n_nodes, int64_t *n_entries)
{
    if (!path || !separator || !*separator || !out || !n_nodes || !n_entries) {
        set_error("read_edge_list: null argument or empty separator");
        return ARCTE_E_ARG;
    }
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error(std::string("read_edge_list: cannot open ") + path + ": " + strerror(errno)); return ARCTE_E_ARG; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); set_error("read_edge_list: fstat failed"); return ARCTE_E_ARG; }
    const size_t size = (size_t)sb.st_size;
    const char *base = nullptr;
    if (size > 0) {
        base = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (base == MAP_FAILED) { close(fd); set_error("read_edge_list: mmap failed"); return ARCTE_E_ARG; }
    }
    close(fd);
    const std::string sep(separator);
    int T = pick_threads(n_threads);
    if ((size_t)T > size / (1 << 16) + 1) T = (int)(size / (1 << 16) + 1);
    // chunk boundaries on line starts
    std::vector<size_t> cut((size_t)T + 1, size);
    cut[0] = 0;
    for (int t = 1; t < T; ++t) {
        size_t p = size / (size_t)T * (size_t)t;
        if (p < cut[(size_t)t - 1]) p = cut[(size_t)t - 1];
        const char *nl = p < size ? (const char *)memchr(base + p, '\n', size - p) : nullptr;
        cut[(size_t)t] = nl ? (size_t)(nl - base) + 1 : size;
    }
    std::vector<ParsedChunk> chunks((size_t)T);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] { parse_range(base + cut[(size_t)t], base + cut[(size_t)t + 1], sep, &chunks[(size_t)t]); });
        for (auto &x : th) x.join();
    }
    if (size > 0) munmap((void *)base, size);
    int64_t line0 = 0, total = 0;
    for (auto &c : chunks) {
        if (c.bad_line >= 0) {
            set_error("read_edge_list: cannot parse line " + std::to_string(line0 + c.bad_line + 1) + " of " + path +
                      ": '" + c.bad_text + "' (expected <int><sep><int><sep><float>)");
            return ARCTE_E_ARG;
        }
        line0 += c.lines;
        total += (int64_t)c.src.size();
    }
    arcte_cuda_edge_list *el = new arcte_cuda_edge_list();
    const size_t reserve = (size_t)total * (undirected ? 2 : 1);
    el->row.reserve(reserve);
    el->col.reserve(reserve);
    el->data.reserve(reserve);
    IdMap ids((size_t)(total / 2 + 1024));  // grows on demand
    for (auto &c : chunks) {
        const size_t m = c.src.size();
        constexpr size_t kAhead = 12;
        for (size_t i = 0; i < m; ++i) {
            if (ids.order.size() * 2 + 2 > ids.slots.size()) {  // grow (the initial size is a guess)
                IdMap bigger(ids.slots.size());
                for (int64_t id : ids.order) bigger.get(id);
                ids = std::move(bigger);
            }
            if (i + kAhead < m) {
                ids.prefetch(c.src[i + kAhead]);
                ids.prefetch(c.dst[i + kAhead]);
            }
            const int64_t s = ids.get(c.src[i]);   // source first, then target (datarw.py:88-93)
            const int64_t t = ids.get(c.dst[i]);
            el->row.push_back(s); el->col.push_back(t); el->data.push_back(c.w[i]);
            if (undirected && s != t) {             // datarw.py:103-107
                el->row.push_back(t); el->col.push_back(s); el->data.push_back(c.w[i]);
            }
        }
        std::vector<int64_t>().swap(c.src);
        std::vector<int64_t>().swap(c.dst);
        std::vector<double>().swap(c.w);
    }
    el->node_ids = std::move(ids.order);
    *n_nodes = (int64_t)el->node_ids.size();
    *n_entries = (int64_t)el->row.size();
    *out = el;
    return ARCTE_OK;
}

int arcte_cuda_io_edge_list_copy(const arcte_cuda_edge_list *el, int64_t *row, int64_t *col, double *data,
                                 int64_t *node_ids)
{
    if (!el) { set_error("edge_list_copy: null handle"); return ARCTE_E_ARG; }
    if (row) memcpy(row, el->row.data(), sizeof(int64_t) * el->row.size());
    if (col) memcpy(col, el->col.data(), sizeof(int64_t) * el->col.size());
    if (data) memcpy(data, el->data.data(), sizeof(double) * el->data.size());
    if (node_ids) memcpy(node_ids, el->node_ids.data(), sizeof(int64_t) * el->node_ids.size());
    return ARCTE_OK;
}

void arcte_cuda_io_edge_list_free(arcte_cuda_edge_list *el) { delete el; }

int arcte_cuda_io_write_features(const char *path, const char *separator, int64_t n_rows, const int64_t *indptr,
                                 const int32_t *indices, const double *data, const int64_t *node_ids,
                                 int n_threads, int64_t *bytes_written)
{
    if (!path || !separator || n_rows < 0 || !indptr || (n_rows > 0 && !node_ids)) {
        set_error("write_features: null argument");
        return ARCTE_E_ARG;
    }
    const int64_t nnz = indptr[n_rows];
    if (nnz > 0 && (!indices || !data)) { set_error("write_features: null arrays"); return ARCTE_E_ARG; }
    const std::string sep(separator);
    const size_t sl = sep.size();
    int T = pick_threads(n_threads);
    if ((int64_t)T > nnz / (1 << 16) + 1) T = (int)(nnz / (1 << 16) + 1);
    // row ranges balanced by stored entries
    std::vector<int64_t> rcut((size_t)T + 1, n_rows);
    rcut[0] = 0;
    for (int t = 1; t < T; ++t) {
        const int64_t target = nnz / T * t;
        rcut[(size_t)t] = std::lower_bound(indptr, indptr + n_rows + 1, target) - indptr;
        if (rcut[(size_t)t] < rcut[(size_t)t - 1]) rcut[(size_t)t] = rcut[(size_t)t - 1];
        if (rcut[(size_t)t] > n_rows) rcut[(size_t)t] = n_rows;
    }
    // pass 1: bytes per range (int(value) truncates toward zero like Python's int(), datarw.py:140)
    std::vector<int64_t> bytes((size_t)T, 0);
    std::atomic<int64_t> bad(-1);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                int64_t b = 0;
                for (int64_t r = rcut[(size_t)t]; r < rcut[(size_t)t + 1]; ++r) {
                    const int64_t cnt = indptr[r + 1] - indptr[r];
                    b += cnt * (int_len(node_ids[r]) + 2 * (int64_t)sl + 1);
                    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
                        const double v = data[k];
                        if (!(fabs(v) < 9.0e18)) { bad.store(k); return; }
                        b += int_len((int64_t)indices[k]) + int_len((int64_t)v);
                    }
                }
                bytes[(size_t)t] = b;
            });
        for (auto &x : th) x.join();
    }
    if (bad.load() >= 0) {
        set_error("write_features: value at stored entry " + std::to_string(bad.load()) + " is not a finite integer-sized number");
        return ARCTE_E_ARG;
    }
    std::vector<int64_t> off((size_t)T + 1, 0);
    for (int t = 0; t < T; ++t) off[(size_t)t + 1] = off[(size_t)t] + bytes[(size_t)t];
    const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { set_error(std::string("write_features: cannot open ") + path + ": " + strerror(errno)); return ARCTE_E_ARG; }
    // pass 2: every thread formats into a private buffer and pwrite()s it at its own offset
    std::atomic<int> io_err(0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                const size_t cap = (size_t)4 << 20;
                std::vector<char> buf(cap + 256 + 2 * sl);
                char *p = buf.data();
                int64_t pos = off[(size_t)t];
                auto flush = [&]() {
                    size_t left = (size_t)(p - buf.data());
                    const char *q = buf.data();
                    while (left > 0) {
                        const ssize_t w = pwrite(fd, q, left, (off_t)pos);
                        if (w < 0) { if (errno == EINTR) continue; io_err.store(errno); return; }
                        q += w; left -= (size_t)w; pos += w;
                    }
                    p = buf.data();
                };
                for (int64_t r = rcut[(size_t)t]; r < rcut[(size_t)t + 1] && !io_err.load(); ++r) {
                    char head[32];
                    const size_t hl = (size_t)(put_int(head, node_ids[r]) - head);
                    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) {
                        memcpy(p, head, hl); p += hl;
                        memcpy(p, sep.data(), sl); p += sl;
                        p = put_int(p, (int64_t)indices[k]);
                        memcpy(p, sep.data(), sl); p += sl;
                        p = put_int(p, (int64_t)data[k]);
                        *p++ = '\n';
                        if ((size_t)(p - buf.data()) >= cap) flush();
                    }
                }
                flush();
            });
        for (auto &x : th) x.join();
    }
    const int rc = close(fd);
    if (io_err.load() || rc != 0) {
        set_error(std::string("write_features: write failed: ") + strerror(io_err.load() ? io_err.load() : errno));
        return ARCTE_E_ARG;
    }
    if (bytes_written) *bytes_written = off[(size_t)T];
    return ARCTE_OK;
}


// Fills count doubles with `value` using all host threads (hostmem.py keeps pooled page-locked
// blocks pre-filled with 1.0 so that the value array of a feature matrix need not be copied).
int arcte_cuda_host_fill_f64(double *p, int64_t count, double value, int n_threads)
{
    if (!p || count < 0) { set_error("host_fill: bad arguments"); return ARCTE_E_ARG; }
    int T = pick_threads(n_threads);
    if ((int64_t)T > count / (1 << 20) + 1) T = (int)(count / (1 << 20) + 1);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([=] {
            const int64_t b = count / T * t, e = t == T - 1 ? count : count / T * (t + 1);
            for (int64_t i = b; i < e; ++i) p[i] = value;
        });
    for (auto &x : th) x.join();
    return ARCTE_OK;
}

}  // extern "C"
