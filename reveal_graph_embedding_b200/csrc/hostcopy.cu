// hostcopy.cu -- large copies between the device and ORDINARY (pageable) host memory at PCIe rate.
//
// The caller of arcte() hands over scipy arrays and gets scipy arrays back (arcte.py:591-688): plain
// numpy memory.  A cudaMemcpy into pageable memory is staged by the driver through one small pinned
// buffer at a few GB/s, and page-locking gigabytes per call (cudaHostRegister / cudaHostAlloc) costs more
// than the copy it speeds up.  Instead the context owns a small fixed ring of pinned slots (2 per worker
// thread, 4 MB each): every worker thread streams its share of the chunks device -> pinned slot with
// cudaMemcpyAsync on its own stream and copies the previous slot into the caller's array while the next
// chunk is in flight.  The first touch of the destination pages (a fresh numpy array is untouched
// memory) is thereby spread over all the threads as well, and the range is advised to use huge pages.
#include <sys/mman.h>
#include <sys/uio.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace arcte {

constexpr size_t kSlotBytes = size_t(4) << 20;
constexpr size_t kStreamedMin = size_t(8) << 20;  // below this one plain copy is as fast

static int host_threads(int want)
{
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int t = want > 0 ? want : hw;
    if (t > hw) t = hw;
    if (t > kMaxCopyThreads) t = kMaxCopyThreads;
    return t;
}

static void advise_huge(void *p, size_t bytes)
{
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    uintptr_t lo = ((uintptr_t)p + page - 1) & ~(uintptr_t)(page - 1);
    uintptr_t hi = ((uintptr_t)p + bytes) & ~(uintptr_t)(page - 1);
    if (hi > lo) (void)madvise((void *)lo, hi - lo, MADV_HUGEPAGE);  // best effort
}

// Slots are allocated per copy thread and never re-allocated: growing the ring costs only the new threads' slots
// (page-locking is about 0.5 ms per MB, and the first call of a process grows the ring as its copies get larger).
static int ensure_ring(arcte_cuda_ctx *c, int threads)
{
    HostRing &r = c->ring;
    std::lock_guard<std::mutex> lock(r.mutex);
    if (r.n_threads >= threads) return ARCTE_OK;
    const auto t0 = std::chrono::steady_clock::now();
    const int before = r.n_threads;
    for (int t = r.n_threads; t < threads; ++t) {
        ARCTE_CUDA_TRY(cudaHostAlloc(&r.pinned[t], 2 * kSlotBytes, cudaHostAllocPortable));
        ARCTE_CUDA_TRY(cudaStreamCreateWithFlags(&r.streams[t], cudaStreamNonBlocking));
        ARCTE_CUDA_TRY(cudaEventCreateWithFlags(&r.events[2 * t], cudaEventDisableTiming));
        ARCTE_CUDA_TRY(cudaEventCreateWithFlags(&r.events[2 * t + 1], cudaEventDisableTiming));
        r.n_threads = t + 1;
    }
    if (getenv("ARCTE_CUDA_DEBUG"))
        fprintf(stderr, "[arcte] pinned ring %d -> %d threads x 2 x 4 MB: %.1f ms\n", before, threads,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return ARCTE_OK;
}

void free_ring(arcte_cuda_ctx *c)
{
    HostRing &r = c->ring;
    std::lock_guard<std::mutex> lock(r.mutex);
    for (int t = 0; t < r.n_threads; ++t) {
        cudaStreamDestroy(r.streams[t]);
        cudaEventDestroy(r.events[2 * t]);
        cudaEventDestroy(r.events[2 * t + 1]);
        if (r.pinned[t]) cudaFreeHost(r.pinned[t]);
        r.pinned[t] = nullptr;
    }
    r.n_threads = 0;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// One job = one contiguous range.  Several jobs run in the same pool of threads (chunks of all jobs are
// dealt round-robin), host-side fills included, so that a copy and a fill overlap.
struct CopyJob {
    char *host;
    char *dev;
    size_t bytes;
    bool to_host;
};

// Destination process of a device -> host job: 0 = this one (memcpy); otherwise the host addresses are virtual
// addresses of THAT process and the pinned slot is written there with process_vm_writev (one process per GPU:
// every rank streams its rows straight into the result arrays of rank 0 through its own PCIe link).
static bool put(pid_t pid, void *dst, const void *src, size_t len)
{
    if (pid == 0) {
        memcpy(dst, src, len);
        return true;
    }
    size_t done = 0;
    while (done < len) {
        struct iovec l = {(char *)const_cast<void *>(src) + done, len - done}, r = {(char *)dst + done, len - done};
        const ssize_t k = process_vm_writev(pid, &l, 1, &r, 1, 0);
        if (k <= 0) return false;
        done += (size_t)k;
    }
    return true;
}

static int run_jobs(arcte_cuda_ctx *c, const std::vector<CopyJob> &jobs, double *fill, size_t fill_count, double fill_value,
                    int n_threads, pid_t pid = 0)
{
    struct Chunk { int job; size_t off, len; };   // job == -1: fill
    std::vector<Chunk> copies, fills, chunks;
    for (size_t j = 0; j < jobs.size(); ++j)
        for (size_t off = 0; off < jobs[j].bytes; off += kSlotBytes)
            copies.push_back({(int)j, off, std::min(kSlotBytes, jobs[j].bytes - off)});
    const size_t fill_bytes = fill_count * sizeof(double);
    for (size_t off = 0; off < fill_bytes; off += kSlotBytes) fills.push_back({-1, off, std::min(kSlotBytes, fill_bytes - off)});
    // copy and fill chunks interleaved in proportion, so that the fills run while the PCIe copies are in flight
    for (size_t a = 0, b = 0; a < copies.size() || b < fills.size();) {
        if (b >= fills.size() || (a < copies.size() && a * fills.size() <= b * copies.size())) chunks.push_back(copies[a++]);
        else chunks.push_back(fills[b++]);
    }
    if (chunks.empty()) return ARCTE_OK;
    int T = host_threads(n_threads);
    if ((size_t)T > chunks.size()) T = (int)chunks.size();
    ARCTE_TRY(ensure_ring(c, T));
    if (pid == 0) {
        for (const CopyJob &j : jobs)
            if (j.to_host) advise_huge(j.host, j.bytes);
        if (fill) advise_huge(fill, fill_bytes);
    }
    std::vector<double> fill_src;   // remote fills copy from a local block of the value
    if (pid != 0 && fill) fill_src.assign(kSlotBytes / sizeof(double), fill_value);
    // interleave copy and fill chunks so that every thread sees both kinds
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    HostRing &r = c->ring;
    auto worker = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed = 1; return; }
        char *slot[2] = {(char *)r.pinned[t], (char *)r.pinned[t] + kSlotBytes};
        cudaStream_t st = r.streams[t];
        int k = 0;
        Chunk pending{};   // device -> host chunk whose slot still has to be copied out
        int pending_slot = -1;
        auto drain = [&]() {
            if (pending_slot < 0) return;
            if (cudaEventSynchronize(r.events[2 * t + pending_slot]) != cudaSuccess) failed = 1;
            if (!put(pid, jobs[pending.job].host + pending.off, slot[pending_slot], pending.len)) failed = 2;
            pending_slot = -1;
        };
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= chunks.size() || failed) break;
            const Chunk ch = chunks[i];
            if (ch.job < 0) {
                double *p = (double *)((char *)fill + ch.off);
                if (pid == 0) std::fill(p, p + ch.len / sizeof(double), fill_value);
                else if (!put(pid, p, fill_src.data(), ch.len)) failed = 2;
                continue;
            }
            const CopyJob &jb = jobs[ch.job];
            const int s = k & 1;
            ++k;
            if (jb.to_host) {
                if (pending_slot == s) drain();
                if (cudaMemcpyAsync(slot[s], jb.dev + ch.off, ch.len, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaEventRecord(r.events[2 * t + s], st) != cudaSuccess) {
                    failed = 1;
                    break;
                }
                const Chunk prev = pending;
                const int prev_slot = pending_slot;
                pending = ch;
                pending_slot = s;
                if (prev_slot >= 0 && prev_slot != s) {   // copy the previous chunk out while this one is in flight
                    if (cudaEventSynchronize(r.events[2 * t + prev_slot]) != cudaSuccess) failed = 1;
                    if (!put(pid, jobs[prev.job].host + prev.off, slot[prev_slot], prev.len)) failed = 2;
                }
            } else {
                drain();
                // the slot must not be overwritten while an earlier upload still reads it
                if (cudaEventSynchronize(r.events[2 * t + s]) != cudaSuccess) failed = 1;
                memcpy(slot[s], jb.host + ch.off, ch.len);
                if (cudaMemcpyAsync(jb.dev + ch.off, slot[s], ch.len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaEventRecord(r.events[2 * t + s], st) != cudaSuccess) {
                    failed = 1;
                    break;
                }
            }
        }
        drain();
        if (cudaStreamSynchronize(st) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (std::thread &th : pool) th.join();
    if (failed) {
        (void)cudaGetLastError();
        set_error(failed == 2 ? std::string("streamed host copy: process_vm_writev into process ") + std::to_string((long)pid) +
                                    " failed (" + strerror(errno) + ")"
                              : std::string("streamed host copy failed"));
        return ARCTE_E_CUDA;
    }
    return ARCTE_OK;
}

// Device -> host.  Everything queued on the context's stream is finished first.
int copy_to_host(arcte_cuda_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return ARCTE_OK;
    if (bytes < kStreamedMin || is_pinned(dst)) {
        ARCTE_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return ARCTE_OK;
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<CopyJob> jobs{{(char *)dst, (char *)const_cast<void *>(src), bytes, true}};
    return run_jobs(c, jobs, nullptr, 0, 0.0, 0);
}

// Host -> device; returns when the data is on the device.
int copy_from_host(arcte_cuda_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return ARCTE_OK;
    if (bytes < kStreamedMin || is_pinned(src)) {
        ARCTE_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return ARCTE_OK;
    }
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<CopyJob> jobs{{(char *)const_cast<void *>(src), (char *)dst, bytes, false}};
    return run_jobs(c, jobs, nullptr, 0, 0.0, 0);
}

// The assembled feature block to the host in one pass of the thread pool: row pointers and column indices
// are copied; the values are copied too, or -- when the caller knows they are structural (all 1.0 except
// the self-loop diagonals it patches itself, arcte.py:379-381, :676-679) -- written as ones by the same
// threads while the indices stream in, which saves two thirds of the PCIe bytes.
int fetch_features(arcte_cuda_ctx *c, int64_t *host_indptr, int32_t *host_indices, double *host_data, int values_are_ones,
                   int n_threads, long pid)
{
    ARCTE_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<CopyJob> jobs;
    const size_t nnz = (size_t)c->out_nnz;
    if (host_indptr) jobs.push_back({(char *)host_indptr, c->out_indptr.as<char>(), sizeof(int64_t) * (size_t)(c->out_rows + 1), true});
    if (host_indices && nnz) jobs.push_back({(char *)host_indices, c->out_indices.as<char>(), sizeof(int32_t) * nnz, true});
    double *fill = nullptr;
    if (host_data && nnz) {
        if (values_are_ones) fill = host_data;
        else jobs.push_back({(char *)host_data, c->out_data.as<char>(), sizeof(double) * nnz, true});
    }
    return run_jobs(c, jobs, fill, fill ? nnz : 0, 1.0, n_threads, (pid_t)(pid == (long)getpid() ? 0 : pid));
}

int host_write_remote(long pid, void *remote_dst, const void *local_src, size_t bytes)
{
    if (!put((pid_t)(pid == (long)getpid() ? 0 : pid), remote_dst, local_src, bytes)) {
        set_error(std::string("process_vm_writev into process ") + std::to_string(pid) + " failed (" + strerror(errno) + ")");
        return ARCTE_E_ARG;
    }
    return ARCTE_OK;
}

void host_advise_huge(void *p, size_t bytes) { advise_huge(p, bytes); }

// ---- a value array of ones that costs no fill ------------------------------------------------------------
// Every stored value of the feature matrix is 1.0 (arcte.py:379-381, :676-679) except a handful of 2.0 diagonals,
// and writing 4 GB of ones into fresh pages was two thirds of the host side of a call.  Instead one 32 MB
// in-memory file (memfd) is filled with 1.0 once per process and mapped MAP_PRIVATE over and over, back to back,
// to cover the array: the pages are shared with the file until somebody writes to them (copy-on-write), so the
// array is ordinary writable memory to the caller and costs a few hundred mmap calls to create.
constexpr size_t kOnesTemplate = size_t(32) << 20;
static int g_ones_fd = -1;
static std::mutex g_ones_mutex;

int host_ones_alloc(size_t count, void **out, size_t *mapped_bytes)
{
    *out = nullptr;
    *mapped_bytes = 0;
    {
        std::lock_guard<std::mutex> lock(g_ones_mutex);
        if (g_ones_fd < 0) {
            const int fd = memfd_create("arcte_ones", MFD_CLOEXEC);
            if (fd < 0 || ftruncate(fd, (off_t)kOnesTemplate) != 0) {
                if (fd >= 0) close(fd);
                set_error(std::string("ones template: memfd_create/ftruncate failed (") + strerror(errno) + ")");
                return ARCTE_E_ARG;
            }
            void *t = mmap(nullptr, kOnesTemplate, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            if (t == MAP_FAILED) {
                close(fd);
                set_error(std::string("ones template: mmap failed (") + strerror(errno) + ")");
                return ARCTE_E_ARG;
            }
            std::fill((double *)t, (double *)t + kOnesTemplate / sizeof(double), 1.0);
            munmap(t, kOnesTemplate);
            g_ones_fd = fd;
        }
    }
    const size_t bytes = ((count * sizeof(double) + kOnesTemplate - 1) / kOnesTemplate) * kOnesTemplate;
    if (bytes == 0) return ARCTE_OK;
    char *base = (char *)mmap(nullptr, bytes, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == MAP_FAILED) {
        set_error(std::string("ones array: address range reservation failed (") + strerror(errno) + ")");
        return ARCTE_E_NOMEM;
    }
    for (size_t off = 0; off < bytes; off += kOnesTemplate) {
        if (mmap(base + off, kOnesTemplate, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_FIXED, g_ones_fd, 0) == MAP_FAILED) {
            munmap(base, bytes);
            set_error(std::string("ones array: mapping the template failed (") + strerror(errno) + ")");
            return ARCTE_E_NOMEM;
        }
    }
    *out = base;
    *mapped_bytes = bytes;
    return ARCTE_OK;
}

void host_ones_free(void *p, size_t mapped_bytes)
{
    if (p && mapped_bytes) munmap(p, mapped_bytes);
}

}  // namespace arcte
