// az_capi.cu -- extern "C" boundary (include/astroz_b200.h).  Host orchestration: the device branch of
// src/Constellation.zig (init :101-200, propagate :245-308, propagateConstellation :541-605) and the
// single-satellite exports of src/c_api/sgp4.zig.  No CPU propagation path exists in this library:
// without a CUDA device every propagate call returns ASTROZ_NO_DEVICE / ASTROZ_CUDA_ERROR.
#include "../../include/astroz_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <map>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "az_ingest.cuh"
#include "az_kernels.cuh"
#include "az_tables.hpp"

namespace {

thread_local std::string g_lastError;
std::mutex g_blockMutex;
std::map<void *, size_t> g_blocks;  // blocks made by astroz_cuda_constellation_host_block: host_free unregisters + unmaps

int32_t cuda_fail(cudaError_t e, const char *what) {
    g_lastError = std::string(what) + ": " + cudaGetErrorString(e);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? ASTROZ_NO_DEVICE : ASTROZ_CUDA_ERROR;
}
#define AZ_CUDA(expr)                                        \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);  \
    } while (0)

#define AZ_SINGLE(c)                                                                                        \
    do {                                                                                                    \
        if ((c) && (c)->multi()) {                                                                          \
            g_lastError = "this entry point works on device pointers of ONE GPU: create the handle with "  \
                          "device >= 0 (a device = -1 handle spans several GPUs)";                          \
            return ASTROZ_VALUE_ERROR;                                                                      \
        }                                                                                                   \
    } while (0)

int32_t status_to_code(int st) {  // kernel-level code -> C API code (src/c_api/sgp4.zig:22-28)
    switch (st) {
        case az::kOk: return ASTROZ_OK;
        case az::kDecayed: return ASTROZ_DECAYED;
        case az::kInvalidEcc: return ASTROZ_INVALID_ECC;
        case az::kDeepSpace: return ASTROZ_DEEP_SPACE;
        case az::kOom: return ASTROZ_ALLOC_FAILED;
        case az::kBadTle: return ASTROZ_BAD_TLE_LENGTH;
        default: return ASTROZ_UNKNOWN;
    }
}

// ---- delivery into caller-owned PAGEABLE host memory --------------------------------------------------------------
// The reference writes into whatever slices the caller hands in (src/Constellation.zig:245-258; the Python layer
// passes plain numpy buffers, bindings/python/src/satrec.zig:917-942).  A device->host DMA into pageable memory is
// staged by the driver through a small internal buffer, synchronously, at a fraction of the PCIe rate.  Instead the
// result leaves the GPU in pieces into a ring of pinned slots owned by the handle (full-rate DMA), and a small pool of
// host threads copies each landed piece to its final place while the next pieces are in flight.
// Piece size: large enough that (a) the wake-up of the copy threads is amortised and (b) each thread's share (a few MB)
// is above libc's non-temporal threshold, so the destination lines are streamed instead of read for ownership first.
constexpr size_t kPieceBytes = 32u << 20;
constexpr int kRingSlots = 3;

// Streaming copy for the landed pieces: the destination is written once and not read again by this library, so the
// stores bypass the cache (no read-for-ownership of the destination lines, no eviction of the caller's working set):
// a third less memory traffic per byte than a cached copy.  AVX2 hosts; anything else uses memcpy.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void stream_copy_avx2(char *dst, const char *src, size_t n) {
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
    if (head) {
        const size_t h = std::min(head, n);
        std::memcpy(dst, src, h);
        dst += h; src += h; n -= h;
    }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) std::memcpy(dst + i, src + i, n - i);
}
static void stream_copy(char *dst, const char *src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= (64u << 10)) stream_copy_avx2(dst, src, n);
    else std::memcpy(dst, src, n);
}
#else
static void stream_copy(char *dst, const char *src, size_t n) { std::memcpy(dst, src, n); }
#endif

class CopyPool {  // process-wide, created on first use, never destroyed (workers sleep on the condition variable)
public:
    static CopyPool &get() {
        static CopyPool *p = new CopyPool();
        return *p;
    }
    int threads() const { return (int)workers_.size(); }
    // copy `rows` rows of rowBytes from a contiguous source to a destination with pitch hpitch, split over the pool
    void copy(char *dst, const char *src, size_t rows, size_t rowBytes, size_t hpitch) {
        const size_t total = rows * rowBytes;
        const int parts = (int)std::max<size_t>(1, std::min<size_t>(workers_.size(), total / (1u << 20)));
        if (parts <= 1 || workers_.empty()) {
            run(dst, src, 0, rows, rowBytes, hpitch, 0, total);
            return;
        }
        std::atomic<int> left(parts);
        std::mutex dm;
        std::condition_variable dcv;
        for (int k = 0; k < parts; ++k) {
            const size_t b0 = total * k / parts, b1 = total * (k + 1) / parts;
            push([=, &left, &dm, &dcv] {
                run(dst, src, 0, rows, rowBytes, hpitch, b0, b1);
                if (left.fetch_sub(1) == 1) {
                    std::lock_guard<std::mutex> g(dm);
                    dcv.notify_one();
                }
            });
        }
        std::unique_lock<std::mutex> g(dm);
        dcv.wait(g, [&] { return left.load() == 0; });
    }

private:
    CopyPool() {
        int n = 12;
        if (const char *v = std::getenv("ASTROZ_COPY_THREADS")) n = std::max(0, std::min(64, std::atoi(v)));
        const unsigned hw = std::thread::hardware_concurrency();
        if (hw && (unsigned)n > hw) n = (int)hw;
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
        for (auto &t : workers_) t.detach();
    }
    // bytes [b0, b1) of the logical contiguous source, scattered to rows of the destination
    static void run(char *dst, const char *src, size_t, size_t, size_t rowBytes, size_t hpitch, size_t b0, size_t b1) {
        if (hpitch == rowBytes) {
            stream_copy(dst + b0, src + b0, b1 - b0);
            return;
        }
        size_t b = b0;
        while (b < b1) {
            const size_t r = b / rowBytes, o = b % rowBytes;
            const size_t len = std::min(rowBytes - o, b1 - b);
            stream_copy(dst + r * hpitch + o, src + b, len);
            b += len;
        }
    }
    void push(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return !q_.empty(); });
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
};

struct Piece {  // one ring-sized piece of a deferred delivery: rows x rowBytes, contiguous on the device
    const char *dsrc;
    char *hdst;
    size_t rows, rowBytes, hpitch;
    int chunk;  // the grid chunk whose kernels produce it (chunkDone[chunk])
};

bool is_pageable(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

template <typename T>
struct DevBuf {  // grow-only device buffer
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Worker threads of a multi-device handle: one per shard after the first (the caller's thread serves shard 0), parked on
// a condition variable between calls.  Creating eight std::threads per call instead put the last shard's first launch
// ~0.2 ms behind the first one's -- 6 % of a 3.3 ms call at eight GPUs.
struct ShardWorkers {
    std::mutex m;
    std::condition_variable wake, done;
    std::vector<std::thread> threads;
    const std::function<int32_t(size_t)> *work = nullptr;
    std::vector<int32_t> rc;
    std::vector<std::string> err;
    uint64_t generation = 0;
    size_t pending = 0;
    bool stop = false;

    void start(size_t nShards);
    int32_t run(const std::function<int32_t(size_t)> &w);
    ~ShardWorkers() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        wake.notify_all();
        for (auto &t : threads) t.join();
    }
};

struct Constellation {
    int device = 0;
    az::CatalogTables cat;
    az::GravConsts g{};
    double sdp4EpochMin = INFINITY, sdp4EpochMax = -INFINITY;  // epoch span of the deep-space records
    cudaStream_t stream = nullptr, copyStream = nullptr, auxStream = nullptr;  // aux: the deep-space grid of a mixed call
    cudaEvent_t forkEv = nullptr, joinEv = nullptr;
    // element tables (resident for the life of the handle)
    DevBuf<double> dTiles, dToff;
    DevBuf<uint32_t> dSgp4Orig, dSdp4Orig, dIdentity, dSdp4Identity;
    DevBuf<az::Sdp4Sat> dSdp4;
    // per-call time axis: tbase | jdFull | gsin | gcos
    // host staging rotates over kSlots pinned buffers so the host can queue several calls ahead of the GPU
    // without stalling on the previous call's asynchronous upload
    static constexpr int kSlots = 4;
    DevBuf<double> dTime;
    double *hTimeSlot[kSlots] = {};
    size_t hTimeSlotCap[kSlots] = {};
    cudaEvent_t slotCopied[kSlots] = {};
    bool slotPending[kSlots] = {};
    int slot = 0;
    double *hTime = nullptr;          // the slot in use by the current call
    cudaEvent_t timeCopied = nullptr;  // its event
    bool timePending = false;          // kept for call sites: set after recording timeCopied
    // the time axis last uploaded by upload_time_axis: a repeated call with the same jd/fr (a propagation loop over a
    // fixed grid) reuses the device copy instead of re-uploading it
    std::vector<double> cachedJd, cachedFr;
    bool cachedGmst = false, cacheValid = false;
    double cachedRef = 0.0, cachedJdMin = 0.0, cachedJdMax = 0.0;
    cudaEvent_t axisReady = nullptr;   // recorded after the cached axis' upload
    cudaStream_t axisStream = nullptr;
    // stateless-path epoch offsets
    DevBuf<double> dToffCall;
    DevBuf<uint8_t> dMask;
    double *hToffCall = nullptr;
    size_t hToffCap = 0;
    cudaEvent_t toffCopied = nullptr;
    bool toffPending = false;
    // resonance lattice (depends on the elements only; grown when a call reaches further in time)
    DevBuf<double2> dLattice;
    int latticeNodes = 0;
    // staging for the host-buffer API
    DevBuf<double> dPos, dVel;
    // coarse-screen scratch: hash heads / chains for one batch of epochs, hit buffers, counter
    DevBuf<uint32_t> dHead, dNext, dPairs, dTIdx;
    DevBuf<unsigned long long> dCount;
    // kernel timing
    cudaEvent_t ev[6] = {};  // K1 start/end, K2 start/end, whole call start/end
    cudaEvent_t chunkDone[64] = {};
    bool timed = false, spanTimed = false;
    bool timing = false;  // kernel-time events are recorded only on request (astroz_cuda_constellation_set_timing): six
                          // timed event records per call cost ~12 us of stream time, 20 % of a 1/8-catalog step
    int variant = -1;  // -1 = shipped default; >= 0 selects a tuning variant (ASTROZ_SGP4_VARIANT)
    int chunks = 8;
    // Multi-device handle (device = -1 at creation): the catalog is cut into contiguous satellite ranges, one
    // single-device shard each (the analogue of the reference's thread fan-out, src/Constellation.zig:327-385, with
    // ASTROZ_DEVICES in the role of ASTROZ_THREADS, :61-74).  The top-level object then holds only the catalog.
    std::vector<Constellation *> shards;
    std::vector<uint32_t> shardRow0;   // first catalog row of each shard, plus the end (size = shards + 1)
    std::vector<uint32_t> shardNear0;  // first near-earth index of each shard, plus the end
    std::vector<uint32_t> shardDeep0;  // first deep-space index of each shard, plus the end
    DevBuf<double> dFullPos, dFullVel; // per shard: the WHOLE block, for the replicated (all-gather) propagate
    ShardWorkers *workers = nullptr;   // multi-device handle: parked host threads, one per shard after the first
    bool multi() const { return !shards.empty(); }
    // deferred delivery into pageable host memory (see CopyPool)
    std::vector<Piece> plan;
    char *ring = nullptr;
    cudaEvent_t ringEv[kRingSlots] = {};

    ~Constellation() {
        delete workers;  // joins them: no shard is in use after this line
        for (Constellation *sh : shards) delete sh;
        shards.clear();
        if (!stream) return;  // never opened on a device (a Satrec that was only inspected): nothing to release
        cudaSetDevice(device);
        dTiles.release(); dToff.release(); dSgp4Orig.release(); dSdp4Orig.release(); dIdentity.release(); dSdp4Identity.release();
        dSdp4.release(); dTime.release(); dToffCall.release(); dMask.release(); dLattice.release(); dPos.release(); dVel.release();
        dHead.release(); dNext.release(); dPairs.release(); dTIdx.release(); dCount.release();
        dFullPos.release(); dFullVel.release();
        if (ring) cudaFreeHost(ring);
        for (auto &e : ringEv) if (e) cudaEventDestroy(e);
        for (auto &h : hTimeSlot) if (h) cudaFreeHost(h);
        if (hToffCall) cudaFreeHost(hToffCall);
        for (auto &e : slotCopied) if (e) cudaEventDestroy(e);
        if (toffCopied) cudaEventDestroy(toffCopied);
        if (axisReady) cudaEventDestroy(axisReady);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        for (auto &e : chunkDone) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (copyStream) cudaStreamDestroy(copyStream);
        if (auxStream) cudaStreamDestroy(auxStream);
        if (forkEv) cudaEventDestroy(forkEv);
        if (joinEv) cudaEventDestroy(joinEv);
    }
};

int32_t upload_toff(Constellation *c) {  // src/Constellation.zig:153
    const size_t n = c->cat.sgp4Epoch.size();
    if (n == 0) return ASTROZ_OK;
    std::vector<double> off(n);
    for (size_t i = 0; i < n; ++i) off[i] = (c->cat.referenceEpochJd - c->cat.sgp4Epoch[i]) * 1440.0;
    AZ_CUDA(c->dToff.reserve(n));
    AZ_CUDA(cudaMemcpy(c->dToff.p, off.data(), n * 8, cudaMemcpyHostToDevice));
    return ASTROZ_OK;
}

int32_t open_device(Constellation *c, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_lastError = "no CUDA device available (this library has no CPU propagation path)";
        return ASTROZ_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        g_lastError = "device index out of range";
        return ASTROZ_VALUE_ERROR;
    }
    c->device = device;
    AZ_CUDA(cudaSetDevice(device));
    AZ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    AZ_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    AZ_CUDA(cudaStreamCreateWithFlags(&c->auxStream, cudaStreamNonBlocking));
    AZ_CUDA(cudaEventCreateWithFlags(&c->forkEv, cudaEventDisableTiming));
    AZ_CUDA(cudaEventCreateWithFlags(&c->joinEv, cudaEventDisableTiming));
    for (auto &e : c->slotCopied) AZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    AZ_CUDA(cudaEventCreateWithFlags(&c->toffCopied, cudaEventDisableTiming));
    AZ_CUDA(cudaEventCreateWithFlags(&c->axisReady, cudaEventDisableTiming));
    c->timeCopied = c->slotCopied[0];
    for (auto &ev : c->ev) AZ_CUDA(cudaEventCreate(&ev));
    for (auto &ev : c->chunkDone) AZ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (const char *v = std::getenv("ASTROZ_SGP4_VARIANT")) c->variant = std::atoi(v);
    if (const char *v = std::getenv("ASTROZ_SDP4_VARIANT")) az::set_sdp4_variant(std::atoi(v));
    if (const char *v = std::getenv("ASTROZ_TIMING")) c->timing = std::atoi(v) != 0;
    if (const char *v = std::getenv("ASTROZ_K1_STRIPE")) az::set_sgp4_stripe((uint32_t)std::max(0, std::atoi(v)));
    if (const char *v = std::getenv("ASTROZ_D2H_CHUNKS")) c->chunks = std::max(1, std::min(64, std::atoi(v)));
    return ASTROZ_OK;
}

int32_t finish_create(Constellation *c, int device) {
    int32_t rc0 = open_device(c, device);
    if (rc0 != ASTROZ_OK) return rc0;
    c->g = az::grav_consts(c->cat.grav);
    const az::CatalogTables &t = c->cat;
    if (t.nSgp4) {
        AZ_CUDA(c->dTiles.reserve(t.sgp4Tiles.size()));
        AZ_CUDA(cudaMemcpy(c->dTiles.p, t.sgp4Tiles.data(), t.sgp4Tiles.size() * 8, cudaMemcpyHostToDevice));
        AZ_CUDA(c->dSgp4Orig.reserve(t.sgp4Orig.size()));
        AZ_CUDA(cudaMemcpy(c->dSgp4Orig.p, t.sgp4Orig.data(), t.sgp4Orig.size() * 4, cudaMemcpyHostToDevice));
        std::vector<uint32_t> ident(t.sgp4Orig.size());
        for (size_t i = 0; i < ident.size(); ++i) ident[i] = (uint32_t)std::min<size_t>(i, t.nSgp4 - 1);
        AZ_CUDA(c->dIdentity.reserve(ident.size()));
        AZ_CUDA(cudaMemcpy(c->dIdentity.p, ident.data(), ident.size() * 4, cudaMemcpyHostToDevice));
        int32_t rc = upload_toff(c);
        if (rc != ASTROZ_OK) return rc;
    }
    if (t.nSdp4) {
        AZ_CUDA(c->dSdp4.reserve(t.nSdp4));
        AZ_CUDA(cudaMemcpy(c->dSdp4.p, t.sdp4.data(), t.nSdp4 * sizeof(az::Sdp4Sat), cudaMemcpyHostToDevice));
        AZ_CUDA(c->dSdp4Orig.reserve(t.nSdp4));
        AZ_CUDA(cudaMemcpy(c->dSdp4Orig.p, t.sdp4Orig.data(), t.nSdp4 * 4, cudaMemcpyHostToDevice));
        for (const az::Sdp4Sat &r : t.sdp4) {
            c->sdp4EpochMin = std::min(c->sdp4EpochMin, r.epochJd);
            c->sdp4EpochMax = std::max(c->sdp4EpochMax, r.epochJd);
        }
    }
    return ASTROZ_OK;
}

// Devices of a device = -1 handle: ASTROZ_DEVICE_LIST="0,2,3" names ordinals explicitly (an ordinal may repeat: several
// shards on one GPU, which is how the sharding is exercised on a one-GPU box); otherwise the first ASTROZ_DEVICES of the
// visible devices (all of them when unset) -- the device-count knob in the role of ASTROZ_THREADS
// (src/Constellation.zig:61-74).
std::vector<int> multi_device_list() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) count = 0;
    std::vector<int> devs;
    if (const char *lst = std::getenv("ASTROZ_DEVICE_LIST")) {
        const char *p = lst;
        while (*p) {
            char *end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            if (v >= 0 && v < count) devs.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
        if (!devs.empty()) return devs;
    }
    int want = count;
    if (const char *v = std::getenv("ASTROZ_DEVICES")) {
        const int k = std::atoi(v);
        if (k >= 1) want = std::min(count, k);
    }
    for (int d = 0; d < want; ++d) devs.push_back(d);
    return devs;
}

// Open the handle on `device` (>= 0), or cut the catalog into one shard per device (device = -1).
int32_t finish_create_any(Constellation *c, int device) {
    if (device >= 0) return finish_create(c, device);
    if (device != -1) {
        g_lastError = "device must be a CUDA ordinal, or -1 for every visible device (ASTROZ_DEVICES caps the count)";
        return ASTROZ_VALUE_ERROR;
    }
    const std::vector<int> devs = multi_device_list();
    const az::CatalogTables &t = c->cat;
    // satellite ranges of equal cost (a deep-space cell costs ~2.4 near-earth cells), cut on multiples of 8 rows so every
    // shard's rows start 64-byte aligned in either layout
    size_t want = std::min<size_t>(devs.size(), std::max<uint32_t>(1u, t.n / 64));
    if (want <= 1) return finish_create(c, devs.empty() ? 0 : devs[0]);
    std::vector<double> prefix(t.n + 1, 0.0);
    for (uint32_t i = 0; i < t.n; ++i) prefix[i + 1] = prefix[i] + (t.classes[i] == 0 ? 1.0 : 2.4);
    std::vector<uint32_t> cut(1, 0u);
    for (size_t k = 1; k < want; ++k) {
        const double target = prefix[t.n] * (double)k / (double)want;
        uint32_t r = (uint32_t)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
        r = std::min(t.n, (r + 4) / 8 * 8);
        if (r > cut.back() && r < t.n) cut.push_back(r);
    }
    cut.push_back(t.n);
    c->g = az::grav_consts(t.grav);
    c->device = devs[0];
    c->shardRow0 = cut;
    c->shardNear0.assign(1, 0u);
    c->shardDeep0.assign(1, 0u);
    for (size_t k = 0; k + 1 < cut.size(); ++k) {
        Constellation *sh = new (std::nothrow) Constellation();
        if (!sh) return ASTROZ_ALLOC_FAILED;
        c->shards.push_back(sh);
        az::slice_catalog(t, cut[k], cut[k + 1], sh->cat);
        c->shardNear0.push_back(c->shardNear0.back() + sh->cat.nSgp4);
        c->shardDeep0.push_back(c->shardDeep0.back() + sh->cat.nSdp4);
        const int32_t rc = finish_create(sh, devs[k % devs.size()]);
        if (rc != ASTROZ_OK) return rc;
    }
    return ASTROZ_OK;
}

// Device-side ingest (K5, az_ingest.cu): the element columns are already in HBM; classification, initialisation and
// the table scatter run there.  The host keeps only what later calls need: counts, epochs, classes, row maps.
int32_t ingest_on_device(Constellation *c, az::IngestArgs a, int grav) {
    cudaStream_t s = c->stream;
    az::CatalogTables &t = c->cat;
    t = az::CatalogTables{};
    t.n = a.n;
    t.grav = az::gravity(grav);
    a.grav = t.grav;
    c->g = az::grav_consts(t.grav);
    if (a.n == 0) return ASTROZ_OK;
    const uint32_t blocks = az::ingest_block_count(a.n);
    DevBuf<uint8_t> flags;
    DevBuf<uint32_t> counts;   // blockNear | blockDeep | totals[2]
    DevBuf<unsigned long long> fail;
    DevBuf<int32_t> classes;
    struct Release {
        DevBuf<uint8_t> &a; DevBuf<uint32_t> &b; DevBuf<unsigned long long> &c; DevBuf<int32_t> &d;
        ~Release() { a.release(); b.release(); c.release(); d.release(); }
    } release{flags, counts, fail, classes};
    AZ_CUDA(flags.reserve(a.n));
    AZ_CUDA(counts.reserve((size_t)blocks * 2 + 2));
    AZ_CUDA(fail.reserve(1));
    AZ_CUDA(classes.reserve(a.n));
    a.flags = flags.p;
    a.blockNear = counts.p;
    a.blockDeep = counts.p + blocks;
    a.totals = counts.p + 2 * (size_t)blocks;
    a.firstFail = fail.p;
    a.classes = classes.p;
    AZ_CUDA(cudaMemsetAsync(fail.p, 0xff, 8, s));
    AZ_CUDA(az::launch_ingest_classify(a, s));
    uint32_t totals[2] = {0, 0};
    unsigned long long firstFail = ~0ull;
    AZ_CUDA(cudaMemcpyAsync(totals, a.totals, 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(&firstFail, fail.p, 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (firstFail != ~0ull) {
        g_lastError = "element set " + std::to_string(firstFail >> 8) + " failed to initialise";
        return status_to_code((int)(firstFail & 0xff));
    }
    t.nSgp4 = totals[0];
    t.nSdp4 = totals[1];
    const uint32_t padded = t.sgp4Padded();
    if (t.nSgp4) {
        AZ_CUDA(c->dTiles.reserve((size_t)t.sgp4Tiles_count() * az::kSgp4TileDoubles));
        AZ_CUDA(c->dSgp4Orig.reserve(padded));
        AZ_CUDA(c->dIdentity.reserve(padded));
    }
    if (t.nSdp4) {
        AZ_CUDA(c->dSdp4.reserve(t.nSdp4));
        AZ_CUDA(c->dSdp4Orig.reserve(t.nSdp4));
    }
    a.tiles = c->dTiles.p;
    a.sgp4Orig = c->dSgp4Orig.p;
    a.identity = c->dIdentity.p;
    a.sdp4 = c->dSdp4.p;
    a.sdp4Orig = c->dSdp4Orig.p;
    AZ_CUDA(az::launch_ingest_build(a, s));
    t.epochs.resize(a.n);
    t.classes.resize(a.n);
    t.sgp4Orig.resize(padded);
    t.sdp4Orig.resize(t.nSdp4);
    AZ_CUDA(cudaMemcpyAsync(t.epochs.data(), a.epochJd, (size_t)a.n * 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(t.classes.data(), classes.p, (size_t)a.n * 4, cudaMemcpyDeviceToHost, s));
    if (padded) AZ_CUDA(cudaMemcpyAsync(t.sgp4Orig.data(), c->dSgp4Orig.p, (size_t)padded * 4, cudaMemcpyDeviceToHost, s));
    if (t.nSdp4) AZ_CUDA(cudaMemcpyAsync(t.sdp4Orig.data(), c->dSdp4Orig.p, (size_t)t.nSdp4 * 4, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(&firstFail, fail.p, 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    if (firstFail != ~0ull) {
        g_lastError = "element set " + std::to_string(firstFail >> 8) + " failed to initialise";
        return status_to_code((int)(firstFail & 0xff));
    }
    t.sgp4Epoch.resize(padded);
    for (uint32_t i = 0; i < padded; ++i) t.sgp4Epoch[i] = t.epochs[t.sgp4Orig[i]];
    if (t.nSgp4) t.referenceEpochJd = t.sgp4Epoch[0];  // src/Constellation.zig:139-140
    for (uint32_t i = 0; i < t.nSdp4; ++i) {
        c->sdp4EpochMin = std::min(c->sdp4EpochMin, t.epochs[t.sdp4Orig[i]]);
        c->sdp4EpochMax = std::max(c->sdp4EpochMax, t.epochs[t.sdp4Orig[i]]);
    }
    return upload_toff(c);
}

// host staging for the time axis; waits for the previous call's async upload before reuse
int32_t reserve_time(Constellation *c, size_t nt) {
    c->cacheValid = false;  // every writer of dTime comes through here (or says so itself)
    // the call that last used the current slot recorded slotCopied[slot] (timePending tells us so)
    if (c->timePending) c->slotPending[c->slot] = true;
    c->timePending = false;
    c->slot = (c->slot + 1) % Constellation::kSlots;
    const int k = c->slot;
    if (c->slotPending[k]) {  // kSlots calls ago: almost always long finished
        AZ_CUDA(cudaEventSynchronize(c->slotCopied[k]));
        c->slotPending[k] = false;
    }
    if (nt * 4 > c->hTimeSlotCap[k]) {
        if (c->hTimeSlot[k]) cudaFreeHost(c->hTimeSlot[k]);
        c->hTimeSlot[k] = nullptr;
        c->hTimeSlotCap[k] = 0;
        AZ_CUDA(cudaMallocHost(&c->hTimeSlot[k], nt * 4 * 8));
        c->hTimeSlotCap[k] = nt * 4;
    }
    c->hTime = c->hTimeSlot[k];
    c->timeCopied = c->slotCopied[k];
    AZ_CUDA(c->dTime.reserve(nt * 4));
    return ASTROZ_OK;
}

int32_t ensure_lattice(Constellation *c, int nodes, cudaStream_t s) {
    nodes = std::min(std::max(nodes, 2), 16384);
    if (nodes <= c->latticeNodes) return ASTROZ_OK;
    nodes = std::max(nodes, 32);
    AZ_CUDA(cudaStreamSynchronize(s));  // a previous launch may still read the old lattice
    AZ_CUDA(c->dLattice.reserve((size_t)c->cat.nSdp4 * 2 * nodes));
    AZ_CUDA(az::launch_sdp4_lattice(c->dSdp4.p, c->cat.nSdp4, c->dLattice.p, nodes, s));
    c->latticeNodes = nodes;
    return ASTROZ_OK;
}

struct Launch {  // one grid pass over (satellite range) x (time range)
    uint32_t tile0 = 0, tileCount = 0;  // near-earth tiles
    bool deepSpace = true;              // include the deep-space satellites
    uint32_t t0 = 0, nt = 0;            // epoch range
};

// Queue the kernels for `L` on stream s.  Time arrays must already be on the device.
struct GatherTargets {  // fused all-gather destinations (see astroz_cuda_constellation_propagate_gather)
    int kind = 0;      // 0 none, 1 multicast, 2 peer stores
    int nPeers = 0;
    double *mcPos = nullptr, *mcVel = nullptr;
    double *peerPos[az::kMaxPeers] = {}, *peerVel[az::kMaxPeers] = {};
};

int32_t queue_grid(Constellation *c, const Launch &L, uint32_t ntTotal, double *dPos, double *dVel, uint8_t *dStatus,
                   int mode, int layout, uint32_t outNumSats, uint32_t outSatOffset, cudaStream_t s, bool timeIt,
                   const GatherTargets *gt = nullptr) {
    const az::CatalogTables &t = c->cat;
    const size_t tcap = c->dTime.cap / 4;
    az::GridArgs a;
    a.g = c->g;
    a.nTimes = (layout == 0) ? ntTotal : L.nt;  // satellite-major rows are ntTotal long
    a.outNumSats = outNumSats;
    a.tbase = c->dTime.p + L.t0;
    a.jdFull = c->dTime.p + tcap + L.t0;
    a.gsin = c->dTime.p + 2 * tcap + L.t0;
    a.gcos = c->dTime.p + 3 * tcap + L.t0;
    // outputs: rows are shifted by outSatOffset, epochs by t0
    size_t shift;
    if (layout == 0) shift = ((size_t)outSatOffset * ntTotal + L.t0) * 3;
    else shift = ((size_t)L.t0 * outNumSats + outSatOffset) * 3;
    a.pos = dPos ? dPos + shift : nullptr;
    a.vel = dVel ? dVel + shift : nullptr;
    if (gt && gt->kind) {
        a.gather = gt->kind;
        a.nPeers = gt->nPeers;
        a.mcPos = gt->mcPos ? gt->mcPos + shift : nullptr;
        a.mcVel = gt->mcVel ? gt->mcVel + shift : nullptr;
        for (int p = 0; p < gt->nPeers; ++p) {
            a.peerPos[p] = gt->peerPos[p] ? gt->peerPos[p] + shift : nullptr;
            a.peerVel[p] = gt->peerVel[p] ? gt->peerVel[p] + shift : nullptr;
        }
    }
    a.status = dStatus ? dStatus + (size_t)outSatOffset * ntTotal + L.t0 : nullptr;
    if (layout == 0 && L.nt != ntTotal) {
        g_lastError = "internal: satellite-major launches cover the whole time axis";
        return ASTROZ_UNKNOWN;
    }
    timeIt = timeIt && c->timing;
    const bool doK1 = L.tileCount && t.nSgp4, doK2 = L.deepSpace && t.nSdp4;
    // A mixed call runs its two grids side by side: the deep-space grid is small (a few waves of CTAs at lower
    // fp64-pipe utilisation) and goes first, on the auxiliary stream, so the near-earth CTAs fill the SMs as it drains.
    const bool fork = doK1 && doK2;
    cudaStream_t s2 = fork ? c->auxStream : s;
    if (timeIt) {
        if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[4], s));
        c->spanTimed = true;
    }
    if (fork) {
        AZ_CUDA(cudaEventRecord(c->forkEv, s));
        AZ_CUDA(cudaStreamWaitEvent(s2, c->forkEv, 0));
    }
    if (!fork && timeIt) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    if (fork || !doK1) {
        if (timeIt && !fork) AZ_CUDA(cudaEventRecord(c->ev[1], s));
        if (timeIt) AZ_CUDA(cudaEventRecord(c->ev[2], s2));
        if (doK2) {
            az::GridArgs k2 = a;
            k2.sdp4 = c->dSdp4.p;
            k2.orig = c->dSdp4Orig.p;
            k2.nSats = t.nSdp4;
            k2.lattice = c->dLattice.p;
            k2.latticeNodes = c->latticeNodes;
            AZ_CUDA(az::launch_sdp4_grid(k2, mode, layout, s2));
        }
        if (timeIt) AZ_CUDA(cudaEventRecord(c->ev[3], s2));
    }
    if (doK1) {
        if (fork && timeIt) AZ_CUDA(cudaEventRecord(c->ev[0], s));
        az::GridArgs k1 = a;
        k1.sgp4Tiles = c->dTiles.p + (size_t)L.tile0 * az::kSgp4TileDoubles;
        k1.toff = c->dToff.p + (size_t)L.tile0 * az::kTileSats;
        k1.orig = c->dSgp4Orig.p + (size_t)L.tile0 * az::kTileSats;
        const uint32_t first = L.tile0 * az::kTileSats;
        k1.nSats = std::min<uint32_t>(t.nSgp4 - first, L.tileCount * az::kTileSats);
        AZ_CUDA(az::launch_sgp4_grid(k1, mode, layout, s, c->variant));
        if (timeIt) AZ_CUDA(cudaEventRecord(c->ev[1], s));
        if (!fork && timeIt) {
            if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
            if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
        }
    }
    if (fork) {
        AZ_CUDA(cudaEventRecord(c->joinEv, s2));
        AZ_CUDA(cudaStreamWaitEvent(s, c->joinEv, 0));
    }
    if (timeIt) AZ_CUDA(cudaEventRecord(c->ev[5], s));
    return ASTROZ_OK;
}

// Build the time axis on the host exactly as the reference does (src/Constellation.zig:266-284) and
// queue its upload on s.
int32_t upload_time_axis(Constellation *c, const double *jd, const double *fr, uint32_t nt, int mode, cudaStream_t s,
                         double *jdMin, double *jdMax) {
    if (c->cacheValid && c->cachedJd.size() == nt && c->cachedRef == c->cat.referenceEpochJd &&
        (mode == 0 || c->cachedGmst) && std::memcmp(c->cachedJd.data(), jd, (size_t)nt * 8) == 0 &&
        std::memcmp(c->cachedFr.data(), fr, (size_t)nt * 8) == 0) {
        *jdMin = c->cachedJdMin;
        *jdMax = c->cachedJdMax;
        if (s != c->axisStream) AZ_CUDA(cudaStreamWaitEvent(s, c->axisReady, 0));
        return ASTROZ_OK;
    }
    c->cacheValid = false;
    int32_t rc = reserve_time(c, nt);
    if (rc != ASTROZ_OK) return rc;
    const size_t cap = c->dTime.cap / 4;
    double *tb = c->hTime, *jf = c->hTime + nt, *gs = c->hTime + 2 * (size_t)nt, *gc = c->hTime + 3 * (size_t)nt;
    double lo = INFINITY, hi = -INFINITY;
    for (uint32_t t = 0; t < nt; ++t) {
        const double j = jd[t] + fr[t];
        jf[t] = j;
        tb[t] = (j - c->cat.referenceEpochJd) * 1440.0;
        lo = std::min(lo, j);
        hi = std::max(hi, j);
        if (mode != 0) {
            const double gm = az::julian_to_gmst(j);
            gs[t] = std::sin(gm);
            gc[t] = std::cos(gm);
        }
    }
    *jdMin = lo;
    *jdMax = hi;
    const int arrays = (mode != 0) ? 4 : 2;
    for (int k = 0; k < arrays; ++k)
        AZ_CUDA(cudaMemcpyAsync(c->dTime.p + k * cap, c->hTime + (size_t)k * nt, (size_t)nt * 8, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaEventRecord(c->timeCopied, s));
    c->timePending = true;
    AZ_CUDA(cudaEventRecord(c->axisReady, s));
    c->axisStream = s;
    c->cachedJd.assign(jd, jd + nt);
    c->cachedFr.assign(fr, fr + nt);
    c->cachedGmst = (mode != 0);
    c->cachedRef = c->cat.referenceEpochJd;
    c->cachedJdMin = lo;
    c->cachedJdMax = hi;
    c->cacheValid = true;
    return ASTROZ_OK;
}

int32_t prepare_deep_space(Constellation *c, double jdMin, double jdMax, cudaStream_t s) {
    if (c->cat.nSdp4 == 0) return ASTROZ_OK;
    const double eMin = c->sdp4EpochMin, eMax = c->sdp4EpochMax;
    const double reach = std::max(std::fabs((jdMax - eMin) * 1440.0), std::fabs((jdMin - eMax) * 1440.0));
    return ensure_lattice(c, (int)std::floor(reach / az::kStepp) + 2, s);
}

int32_t check_args(Constellation *c, const void *jd, const void *fr, const void *pos, int mode, int layout) {
    if (!c || !jd || !fr || !pos) return ASTROZ_NULL_POINTER;
    if (mode < 0 || mode > 2 || layout < 0 || layout > 1) {
        g_lastError = "invalid output mode / layout";
        return ASTROZ_VALUE_ERROR;
    }
    return ASTROZ_OK;
}

struct Sgp4Single {
    Constellation *c = nullptr;
    int device = 0;
    bool opened = false;  // streams, events and the device tables are created by the first propagation call
    double epochJd = 0;
    bool deep = false;
    double elements[10] = {};  // ecco inclo nodeo argpo mo no_kozai bstar a no_unkozai epochJd
};

}  // namespace

// ================================================================================================
extern "C" {
#pragma GCC visibility push(default)

uint32_t astroz_cuda_version(void) { return (0u << 16) | (1u << 8) | 0u; }

int32_t astroz_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *astroz_cuda_last_error(void) { return g_lastError.c_str(); }

void *astroz_cuda_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    return p;
}
void astroz_cuda_host_free(void *p) {
    if (!p) return;
    size_t mapped = 0;
    {
        std::lock_guard<std::mutex> g(g_blockMutex);
        auto it = g_blocks.find(p);
        if (it != g_blocks.end()) {
            mapped = it->second;
            g_blocks.erase(it);
        }
    }
    if (mapped) {  // a block from astroz_cuda_constellation_host_block
        cudaHostUnregister(p);
        munmap(p, mapped);
        return;
    }
    cudaFreeHost(p);
}

int32_t astroz_cuda_constellation_create(const char *const *line1, const char *const *line2, uint32_t n, int32_t grav,
                                         int32_t device, astroz_constellation_t *out) {
    if (!out || (n && (!line1 || !line2))) return ASTROZ_NULL_POINTER;
    *out = nullptr;
    Constellation *c = new (std::nothrow) Constellation();
    if (!c) return ASTROZ_ALLOC_FAILED;
    int rc = az::build_catalog(line1, line2, n, grav, c->cat);
    if (rc != az::kOk) {
        delete c;
        return status_to_code(rc);
    }
    int32_t e = finish_create_any(c, device);
    if (e != ASTROZ_OK) {
        delete c;
        return e;
    }
    *out = c;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_create_from_text(const char *text, size_t len, int32_t grav, int32_t device,
                                                   astroz_constellation_t *out) {
    if (!text || !out) return ASTROZ_NULL_POINTER;
    std::vector<std::string> l1, l2;
    az::split_tle_text(text, len, l1, l2);
    std::vector<const char *> p1(l1.size()), p2(l2.size());
    for (size_t i = 0; i < l1.size(); ++i) {
        p1[i] = l1[i].c_str();
        p2[i] = l2[i].c_str();
    }
    return astroz_cuda_constellation_create(p1.data(), p2.data(), (uint32_t)l1.size(), grav, device, out);
}

int32_t astroz_cuda_constellation_create_from_elements(const double *epoch_jd, const double *mean_motion_rev_day,
                                                       const double *ecc, const double *incl_deg, const double *raan_deg,
                                                       const double *argp_deg, const double *ma_deg, const double *bstar,
                                                       uint32_t n, int32_t grav, int32_t device,
                                                       astroz_constellation_t *out) {
    if (!out || (n && (!epoch_jd || !mean_motion_rev_day || !ecc || !incl_deg || !raan_deg || !argp_deg || !ma_deg || !bstar)))
        return ASTROZ_NULL_POINTER;
    *out = nullptr;
    std::vector<az::TleRecord> recs(n);
    for (uint32_t i = 0; i < n; ++i) {
        az::TleRecord &t = recs[i];
        t.satnum = i;
        t.epochJd = epoch_jd[i];
        t.revPerDay = mean_motion_rev_day[i];
        t.ecc = ecc[i];
        t.inclDeg = incl_deg[i];
        t.raanDeg = raan_deg[i];
        t.argpDeg = argp_deg[i];
        t.maDeg = ma_deg[i];
        t.bstar = bstar[i];
    }
    Constellation *c = new (std::nothrow) Constellation();
    if (!c) return ASTROZ_ALLOC_FAILED;
    int rc = az::build_catalog_records(recs.data(), n, grav, c->cat);
    if (rc != az::kOk) {
        delete c;
        return status_to_code(rc);
    }
    int32_t e = finish_create_any(c, device);
    if (e != ASTROZ_OK) {
        delete c;
        return e;
    }
    *out = c;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_create_from_elements_device(
    const double *d_epoch_jd, const double *d_mean_motion_rev_day, const double *d_ecc, const double *d_incl_deg,
    const double *d_raan_deg, const double *d_argp_deg, const double *d_ma_deg, const double *d_bstar, uint32_t n,
    int32_t grav, int32_t device, astroz_constellation_t *out) {
    if (!out || (n && (!d_epoch_jd || !d_mean_motion_rev_day || !d_ecc || !d_incl_deg || !d_raan_deg || !d_argp_deg ||
                       !d_ma_deg || !d_bstar)))
        return ASTROZ_NULL_POINTER;
    *out = nullptr;
    Constellation *c = new (std::nothrow) Constellation();
    if (!c) return ASTROZ_ALLOC_FAILED;
    int32_t e = open_device(c, device);
    if (e == ASTROZ_OK) {
        az::IngestArgs a;
        a.epochJd = d_epoch_jd;
        a.revPerDay = d_mean_motion_rev_day;
        a.ecc = d_ecc;
        a.inclDeg = d_incl_deg;
        a.raanDeg = d_raan_deg;
        a.argpDeg = d_argp_deg;
        a.maDeg = d_ma_deg;
        a.bstar = d_bstar;
        a.n = n;
        e = ingest_on_device(c, a, grav);
    }
    if (e != ASTROZ_OK) {
        delete c;
        return e;
    }
    *out = c;
    return ASTROZ_OK;
}

void astroz_cuda_constellation_free(astroz_constellation_t h) { delete static_cast<Constellation *>(h); }

int32_t astroz_cuda_constellation_counts(astroz_constellation_t h, uint32_t *n, uint32_t *ns, uint32_t *nd) {
    if (!h) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    if (n) *n = c->cat.n;
    if (ns) *ns = c->cat.nSgp4;
    if (nd) *nd = c->cat.nSdp4;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_epochs(astroz_constellation_t h, double *epochs) {
    if (!h || !epochs) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    std::memcpy(epochs, c->cat.epochs.data(), c->cat.epochs.size() * 8);
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_classes(astroz_constellation_t h, int32_t *classes) {
    if (!h || !classes) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    std::memcpy(classes, c->cat.classes.data(), c->cat.classes.size() * 4);
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_get_reference_epoch(astroz_constellation_t h, double *jd) {
    if (!h || !jd) return ASTROZ_NULL_POINTER;
    *jd = static_cast<Constellation *>(h)->cat.referenceEpochJd;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_set_reference_epoch(astroz_constellation_t h, double jd) {
    if (!h) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    if (c->multi()) {
        c->cat.referenceEpochJd = jd;
        for (Constellation *sh : c->shards) {
            const int32_t rc = astroz_cuda_constellation_set_reference_epoch(sh, jd);
            if (rc != ASTROZ_OK) return rc;
        }
        return ASTROZ_OK;
    }
    AZ_CUDA(cudaSetDevice(c->device));
    // kernels still reading the old offsets may sit on the handle's stream or on a caller's stream that was given to a
    // _device call: the whole device is drained, this is a cold configuration call
    AZ_CUDA(cudaDeviceSynchronize());
    c->cat.referenceEpochJd = jd;
    c->cacheValid = false;
    return upload_toff(c);
}

int32_t astroz_cuda_constellation_propagate_device(astroz_constellation_t h, const double *jd, const double *fr,
                                                   uint32_t n_times, double *d_pos, double *d_vel, uint8_t *d_status,
                                                   int32_t mode, int32_t layout, uint32_t out_num_sats,
                                                   uint32_t out_sat_offset, void *stream) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    int32_t rc = check_args(c, jd, fr, d_pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    if (n_times == 0 || c->cat.n == 0) return ASTROZ_OK;
    if (out_num_sats < out_sat_offset + c->cat.n) {  // src/Constellation.zig:255-257 reports a short buffer this way
        g_lastError = "output block smaller than numSatellites rows";
        return ASTROZ_DECAYED;
    }
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    double jdMin, jdMax;
    rc = upload_time_axis(c, jd, fr, n_times, mode, s, &jdMin, &jdMax);
    if (rc != ASTROZ_OK) return rc;
    rc = prepare_deep_space(c, jdMin, jdMax, s);
    if (rc != ASTROZ_OK) return rc;
    Launch L;
    L.tileCount = c->cat.sgp4Tiles_count();
    L.nt = n_times;
    rc = queue_grid(c, L, n_times, d_pos, d_vel, d_status, mode, layout, out_num_sats, out_sat_offset, s, true);
    c->timed = (rc == ASTROZ_OK) && c->timing;
    return rc;
}

// ---- deep-space members only (Constellation.propagateSdp4Constellation, src/Constellation.zig:611-674) --------
static int32_t sdp4_into_common(Constellation *c, const double *jd, const double *fr, uint32_t nt, double *dPos,
                                double *dVel, int mode, int layout, uint32_t outNumSats, uint32_t satOffset,
                                cudaStream_t s) {
    const uint32_t nd = c->cat.nSdp4;
    if (c->dSdp4Identity.cap < nd) {
        std::vector<uint32_t> ident(nd);
        for (uint32_t i = 0; i < nd; ++i) ident[i] = i;  // origIndices = sat_offset + i, satrec.zig:628-631
        AZ_CUDA(c->dSdp4Identity.reserve(nd));
        AZ_CUDA(cudaMemcpy(c->dSdp4Identity.p, ident.data(), (size_t)nd * 4, cudaMemcpyHostToDevice));
    }
    double jdMin, jdMax;
    int32_t rc = upload_time_axis(c, jd, fr, nt, mode, s, &jdMin, &jdMax);
    if (rc != ASTROZ_OK) return rc;
    rc = prepare_deep_space(c, jdMin, jdMax, s);
    if (rc != ASTROZ_OK) return rc;
    const size_t tcap = c->dTime.cap / 4;
    az::GridArgs a;
    a.g = c->g;
    a.nTimes = nt;
    a.outNumSats = outNumSats;
    a.jdFull = c->dTime.p + tcap;
    a.gsin = c->dTime.p + 2 * tcap;
    a.gcos = c->dTime.p + 3 * tcap;
    const size_t shift = (layout == 0) ? (size_t)satOffset * nt * 3 : (size_t)satOffset * 3;
    a.pos = dPos + shift;
    a.vel = dVel ? dVel + shift : nullptr;
    a.sdp4 = c->dSdp4.p;
    a.orig = c->dSdp4Identity.p;
    a.nSats = nd;
    a.lattice = c->dLattice.p;
    a.latticeNodes = c->latticeNodes;
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
    AZ_CUDA(az::launch_sdp4_grid(a, mode, layout, s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
    c->timed = c->timing;
    c->spanTimed = false;
    return ASTROZ_OK;
}

static int32_t sdp4_into_check(Constellation *c, uint32_t out_num_sats, uint32_t sat_offset, uint32_t *rows) {
    *rows = out_num_sats ? out_num_sats : c->cat.nSdp4;
    if (*rows < sat_offset + c->cat.nSdp4) {  // src/Constellation.zig:626-628 reports a short buffer this way
        g_lastError = "output block smaller than sat_offset + numSdp4 rows";
        return ASTROZ_DECAYED;
    }
    return ASTROZ_OK;
}

int32_t astroz_cuda_sdp4_propagate_into_device(astroz_constellation_t h, const double *jd, const double *fr,
                                               uint32_t n_times, double *d_pos, double *d_vel, int32_t mode,
                                               int32_t layout, uint32_t out_num_sats, uint32_t sat_offset, void *stream) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    int32_t rc = check_args(c, jd, fr, d_pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    if (n_times == 0 || c->cat.nSdp4 == 0) return ASTROZ_OK;
    uint32_t rows;
    rc = sdp4_into_check(c, out_num_sats, sat_offset, &rows);
    if (rc != ASTROZ_OK) return rc;
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return sdp4_into_common(c, jd, fr, n_times, d_pos, d_vel, mode, layout, rows, sat_offset, s);
}

int32_t astroz_cuda_sdp4_propagate_into(astroz_constellation_t h, const double *jd, const double *fr, uint32_t n_times,
                                        double *pos, double *vel, int32_t mode, int32_t layout, uint32_t out_num_sats,
                                        uint32_t sat_offset) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    int32_t rc = check_args(c, jd, fr, pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    const uint32_t nd = c->cat.nSdp4;
    if (n_times == 0 || nd == 0) return ASTROZ_OK;
    uint32_t rows;
    rc = sdp4_into_check(c, out_num_sats, sat_offset, &rows);
    if (rc != ASTROZ_OK) return rc;
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // the deep-space rows are computed as a dense block on the device and land in the caller's (possibly wider)
    // block with one strided copy; rows that belong to other satellites are never touched
    const size_t dense = (size_t)nd * n_times * 3;
    AZ_CUDA(c->dPos.reserve(dense));
    if (vel) AZ_CUDA(c->dVel.reserve(dense));
    rc = sdp4_into_common(c, jd, fr, n_times, c->dPos.p, vel ? c->dVel.p : nullptr, mode, layout, nd, 0, s);
    if (rc != ASTROZ_OK) return rc;
    for (int which = 0; which < (vel ? 2 : 1); ++which) {
        double *dst = which ? vel : pos;
        const double *src = which ? c->dVel.p : c->dPos.p;
        if (layout == 0) {
            AZ_CUDA(cudaMemcpyAsync(dst + (size_t)sat_offset * n_times * 3, src, dense * 8, cudaMemcpyDeviceToHost, s));
        } else {
            AZ_CUDA(cudaMemcpy2DAsync(dst + (size_t)sat_offset * 3, (size_t)rows * 24, src, (size_t)nd * 24,
                                      (size_t)nd * 24, n_times, cudaMemcpyDeviceToHost, s));
        }
    }
    AZ_CUDA(cudaStreamSynchronize(s));
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_propagate_device_f32(astroz_constellation_t h, const double *jd, const double *fr,
                                                       uint32_t n_times, double *d_pos, double *d_vel, int32_t phase64,
                                                       void *stream) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    if (!c || !jd || !fr || !d_pos || !d_vel) return ASTROZ_NULL_POINTER;
    if (n_times == 0 || c->cat.nSgp4 == 0) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    double jdMin, jdMax;
    int32_t rc = upload_time_axis(c, jd, fr, n_times, ASTROZ_MODE_TEME, s, &jdMin, &jdMax);
    if (rc != ASTROZ_OK) return rc;
    az::GridArgs a;
    a.g = c->g;
    a.sgp4Tiles = c->dTiles.p;
    a.toff = c->dToff.p;
    a.orig = c->dSgp4Orig.p;
    a.nSats = c->cat.nSgp4;
    a.tbase = c->dTime.p;
    a.nTimes = n_times;
    a.pos = d_pos;
    a.vel = d_vel;
    a.outNumSats = c->cat.n;
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    AZ_CUDA(az::launch_sgp4_grid_f32(a, phase64, s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
    c->timed = c->timing;
    c->spanTimed = false;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_propagate_gather(astroz_constellation_t h, const double *jd, const double *fr,
                                                   uint32_t n_times, void *const *peer_pos, void *const *peer_vel,
                                                   uint32_t n_peers, void *mc_pos, void *mc_vel, uint32_t out_num_sats,
                                                   uint32_t out_sat_offset, void *stream) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    if (!c || !jd || !fr) return ASTROZ_NULL_POINTER;
    if (!mc_pos && (!peer_pos || n_peers == 0)) return ASTROZ_NULL_POINTER;
    if (n_peers > (uint32_t)az::kMaxPeers) {
        g_lastError = "at most 8 peers (one NVSwitch domain)";
        return ASTROZ_VALUE_ERROR;
    }
    if (n_times == 0 || c->cat.n == 0) return ASTROZ_OK;
    if (out_num_sats < out_sat_offset + c->cat.n) {
        g_lastError = "output block smaller than numSatellites rows";
        return ASTROZ_DECAYED;
    }
    GatherTargets gt;
    gt.kind = mc_pos ? 1 : 2;
    gt.nPeers = (int)n_peers;
    gt.mcPos = static_cast<double *>(mc_pos);
    gt.mcVel = static_cast<double *>(mc_vel);
    for (uint32_t p = 0; p < n_peers && peer_pos; ++p) {
        gt.peerPos[p] = static_cast<double *>(peer_pos[p]);
        gt.peerVel[p] = peer_vel ? static_cast<double *>(peer_vel[p]) : nullptr;
        if (!gt.peerPos[p]) return ASTROZ_NULL_POINTER;
    }
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    double jdMin, jdMax;
    int32_t rc = upload_time_axis(c, jd, fr, n_times, ASTROZ_MODE_TEME, s, &jdMin, &jdMax);
    if (rc != ASTROZ_OK) return rc;
    rc = prepare_deep_space(c, jdMin, jdMax, s);
    if (rc != ASTROZ_OK) return rc;
    Launch L;
    L.tileCount = c->cat.sgp4Tiles_count();
    L.nt = n_times;
    rc = queue_grid(c, L, n_times, nullptr, nullptr, nullptr, ASTROZ_MODE_TEME, ASTROZ_LAYOUT_SATELLITE_MAJOR,
                    out_num_sats, out_sat_offset, s, true, &gt);
    c->timed = (rc == ASTROZ_OK) && c->timing;
    return rc;
}

// Send `rows` rows of rowBytes (contiguous on the device at dsrc) to the host at hdst with pitch hpitch, after the
// kernels of grid chunk `chunk`.  Pinned / registered destinations get the copy queued on the copy stream right away;
// pageable ones are recorded as ring-sized pieces and delivered by run_ring() in the wait half of the call.
static int32_t deliver(Constellation *c, bool pageable, int chunk, const double *dsrc, double *hdst, size_t rows,
                       size_t rowBytes, size_t hpitch) {
    if (!pageable) {
        if (rows == 1 || hpitch == rowBytes)
            AZ_CUDA(cudaMemcpyAsync(hdst, dsrc, rows * rowBytes, cudaMemcpyDeviceToHost, c->copyStream));
        else
            AZ_CUDA(cudaMemcpy2DAsync(hdst, hpitch, dsrc, rowBytes, rowBytes, rows, cudaMemcpyDeviceToHost, c->copyStream));
        return ASTROZ_OK;
    }
    const char *src = reinterpret_cast<const char *>(dsrc);
    char *dst = reinterpret_cast<char *>(hdst);
    if (rows == 1 || hpitch == rowBytes) {  // one contiguous run: cut by bytes
        const size_t total = rows * rowBytes;
        for (size_t b = 0; b < total; b += kPieceBytes)
            c->plan.push_back(Piece{src + b, dst + b, 1, std::min(kPieceBytes, total - b), std::min(kPieceBytes, total - b), chunk});
    } else if (rowBytes > kPieceBytes) {     // very wide rows: each row cut by bytes
        for (size_t r = 0; r < rows; ++r)
            for (size_t b = 0; b < rowBytes; b += kPieceBytes)
                c->plan.push_back(Piece{src + r * rowBytes + b, dst + r * hpitch + b, 1, std::min(kPieceBytes, rowBytes - b),
                                        std::min(kPieceBytes, rowBytes - b), chunk});
    } else {                                // whole rows per piece
        const size_t per = std::max<size_t>(1, kPieceBytes / rowBytes);
        for (size_t r = 0; r < rows; r += per)
            c->plan.push_back(Piece{src + r * rowBytes, dst + r * hpitch, std::min(per, rows - r), rowBytes, hpitch, chunk});
    }
    return ASTROZ_OK;
}

// Drain the deferred deliveries of one handle: up to kRingSlots pieces in flight over PCIe while the pool copies the
// landed one to its final place.
static int32_t run_ring(Constellation *c) {
    if (c->plan.empty()) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    if (!c->ring) {
        AZ_CUDA(cudaMallocHost(&c->ring, kPieceBytes * kRingSlots));
        for (auto &e : c->ringEv) AZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    CopyPool &pool = CopyPool::get();
    const size_t n = c->plan.size();
    auto issue = [&](size_t i) -> cudaError_t {
        const Piece &p = c->plan[i];
        const int slot = (int)(i % kRingSlots);
        cudaError_t e = cudaStreamWaitEvent(c->copyStream, c->chunkDone[p.chunk], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(c->ring + (size_t)slot * kPieceBytes, p.dsrc, p.rows * p.rowBytes, cudaMemcpyDeviceToHost,
                                c->copyStream);
        if (e == cudaSuccess) e = cudaEventRecord(c->ringEv[slot], c->copyStream);
        return e;
    };
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < std::min<size_t>(kRingSlots, n) && e == cudaSuccess; ++i) e = issue(i);
    for (size_t i = 0; i < n && e == cudaSuccess; ++i) {
        const int slot = (int)(i % kRingSlots);
        e = cudaEventSynchronize(c->ringEv[slot]);
        if (e != cudaSuccess) break;
        const Piece &p = c->plan[i];
        pool.copy(p.hdst, c->ring + (size_t)slot * kPieceBytes, p.rows, p.rowBytes, p.hpitch);
        if (i + kRingSlots < n) e = issue(i + kRingSlots);
    }
    c->plan.clear();
    if (e != cudaSuccess) return cuda_fail(e, "pageable delivery ring");
    return ASTROZ_OK;
}

// Host-buffer propagate, in a queue half and a wait half (deliveries to pageable memory are drained in the latter):
// queue = upload the time axis, launch the grid in chunks, start each chunk's device->host copy as soon as its kernels
// finish; wait = drain the streams.  This handle's rows land at rows [rowOffset, rowOffset + n) of a host block with
// totalRows rows (its own block when rowOffset = 0, totalRows = n).
static int32_t propagate_host_queue(Constellation *c, const double *jd, const double *fr, uint32_t n_times, double *pos,
                                    double *vel, int32_t mode, int32_t layout, uint32_t rowOffset, uint32_t totalRows) {
    const uint32_t n = c->cat.n;
    if (n_times == 0 || n == 0) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    const size_t total = (size_t)n * n_times * 3;
    AZ_CUDA(c->dPos.reserve(total));
    if (vel) AZ_CUDA(c->dVel.reserve(total));
    double *dPos = c->dPos.p, *dVel = vel ? c->dVel.p : nullptr;
    cudaStream_t s = c->stream;
    double jdMin, jdMax;
    int32_t rc = upload_time_axis(c, jd, fr, n_times, mode, s, &jdMin, &jdMax);
    if (rc != ASTROZ_OK) return rc;
    rc = prepare_deep_space(c, jdMin, jdMax, s);
    if (rc != ASTROZ_OK) return rc;

    // Pipeline: the grid is cut into chunks whose output is one contiguous block of the result, so the
    // device->host copy of chunk i (copy stream) overlaps the kernels of chunk i+1 (compute stream).
    const uint32_t tiles = c->cat.sgp4Tiles_count();
    const bool bySat = (layout == 0) && c->cat.nSdp4 == 0;  // near-earth rows are the identity map
    const bool byTime = (layout == 1);
    uint32_t units = bySat ? tiles : (byTime ? n_times : 1);
    uint32_t nChunks = (bySat || byTime) ? std::min<uint32_t>((uint32_t)c->chunks, units) : 1;
    if (total * 8 < (8u << 20)) nChunks = 1;
    uint32_t per = (units + nChunks - 1) / nChunks;
    // a thread's epochs are 32 apart and the kernels start their runs at multiples of 64 / 96 epochs from the start of the
    // launch: time chunks that start on multiples of 192 keep every thread's set of epochs -- and with it the series each
    // cell takes, i.e. every result bit -- the same however the call is chunked (one device or many, any chunk count)
    if (byTime && nChunks > 1) per = (per + 191) / 192 * 192;
    const bool posPageable = is_pageable(pos), velPageable = vel && is_pageable(vel);
    c->plan.clear();
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[4], s));  // whole-call span: first kernel of the first chunk ...
    for (uint32_t k = 0; k < nChunks; ++k) {
        const uint32_t u0 = k * per, u1 = std::min(units, u0 + per);
        if (u0 >= u1) break;
        Launch L;
        size_t off, cnt;
        if (bySat) {
            L.tile0 = u0; L.tileCount = u1 - u0; L.deepSpace = false; L.t0 = 0; L.nt = n_times;
            const size_t r0 = (size_t)u0 * az::kTileSats, r1 = std::min<size_t>(n, (size_t)u1 * az::kTileSats);
            off = r0 * n_times * 3;
            cnt = (r1 - r0) * n_times * 3;
        } else if (byTime) {
            L.tile0 = 0; L.tileCount = tiles; L.deepSpace = true; L.t0 = u0; L.nt = u1 - u0;
            off = (size_t)u0 * n * 3;
            cnt = (size_t)(u1 - u0) * n * 3;
        } else {
            L.tile0 = 0; L.tileCount = tiles; L.deepSpace = true; L.t0 = 0; L.nt = n_times;
            off = 0;
            cnt = total;
        }
        rc = queue_grid(c, L, n_times, dPos, dVel, nullptr, mode, layout, n, 0, s, false);
        if (rc != ASTROZ_OK) return rc;
        AZ_CUDA(cudaEventRecord(c->chunkDone[k], s));
        AZ_CUDA(cudaStreamWaitEvent(c->copyStream, c->chunkDone[k], 0));
        for (int which = 0; which < (vel ? 2 : 1); ++which) {
            double *hdst = which ? vel : pos;
            const double *dsrc = (which ? dVel : dPos) + off;
            const bool pg = which ? velPageable : posPageable;
            if (layout == 0) {  // this handle's rows are one contiguous run of the (possibly wider) host block
                rc = deliver(c, pg, (int)k, dsrc, hdst + (size_t)rowOffset * n_times * 3 + off, 1, cnt * 8, cnt * 8);
            } else if (totalRows == n) {
                rc = deliver(c, pg, (int)k, dsrc, hdst + off, 1, cnt * 8, cnt * 8);
            } else {            // time-major into a wider block: n*24 bytes per epoch at a pitch of totalRows*24
                rc = deliver(c, pg, (int)k, dsrc, hdst + ((size_t)u0 * totalRows + rowOffset) * 3, u1 - u0, (size_t)n * 24,
                             (size_t)totalRows * 24);
            }
            if (rc != ASTROZ_OK) return rc;
        }
    }
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[5], s));  // ... to the last kernel of the last chunk
    // last_kernel_ms after a host-buffer call: ms[1] = the span above (all chunks, copies overlapping); the per-kernel
    // slots repeat it
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
    c->timed = c->timing;
    c->spanTimed = true;
    return ASTROZ_OK;
}

static int32_t propagate_host_wait(Constellation *c) {
    if (!c->stream) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    const int32_t rr = run_ring(c);
    if (rr != ASTROZ_OK) return rr;
    AZ_CUDA(cudaStreamSynchronize(c->copyStream));
    AZ_CUDA(cudaStreamSynchronize(c->stream));
    return ASTROZ_OK;
}

// Run `work(k)` (queue + wait of shard k) for every shard of a multi-device handle, one host thread per shard: the
// launches and copies of the devices are issued side by side instead of one device after the other (for 8 GPUs the
// serial issue of ~40 launches and ~130 copies was ~0.4 ms of a 3 ms call).
void ShardWorkers::start(size_t nShards) {
    rc.assign(nShards, ASTROZ_OK);
    err.assign(nShards, std::string());
    for (size_t k = 1; k < nShards; ++k)
        threads.emplace_back([this, k] {
            uint64_t seen = 0;
            for (;;) {
                const std::function<int32_t(size_t)> *w;
                {
                    std::unique_lock<std::mutex> lk(m);
                    wake.wait(lk, [&] { return stop || generation != seen; });
                    if (stop) return;
                    seen = generation;
                    w = work;
                }
                const int32_t r = (*w)(k);
                std::lock_guard<std::mutex> lk(m);
                rc[k] = r;
                if (r != ASTROZ_OK) err[k] = g_lastError;  // thread-local in the worker: carry it back
                if (--pending == 0) done.notify_one();
            }
        });
}

int32_t ShardWorkers::run(const std::function<int32_t(size_t)> &w) {
    {
        std::lock_guard<std::mutex> lk(m);
        work = &w;
        pending = threads.size();
        ++generation;
    }
    wake.notify_all();
    const int32_t r0 = w(0);
    const std::string e0 = (r0 != ASTROZ_OK) ? g_lastError : std::string();
    std::unique_lock<std::mutex> lk(m);
    done.wait(lk, [&] { return pending == 0; });
    rc[0] = r0;
    err[0] = e0;
    for (size_t k = 0; k < rc.size(); ++k)
        if (rc[k] != ASTROZ_OK) {
            g_lastError = err[k];
            return rc[k];
        }
    return ASTROZ_OK;
}

static int32_t for_each_shard(Constellation *c, const std::function<int32_t(size_t)> &work) {
    if (!c->workers) {  // first call through this handle (calls on one handle are serial, include/astroz_b200.h)
        c->workers = new ShardWorkers();
        c->workers->start(c->shards.size());
    }
    return c->workers->run(work);
}

int32_t astroz_cuda_constellation_propagate(astroz_constellation_t h, const double *jd, const double *fr,
                                            uint32_t n_times, double *pos, double *vel, int32_t mode, int32_t layout) {
    Constellation *c = static_cast<Constellation *>(h);
    int32_t rc = check_args(c, jd, fr, pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    if (n_times == 0 || c->cat.n == 0) return ASTROZ_OK;
    if (!c->multi()) {
        rc = propagate_host_queue(c, jd, fr, n_times, pos, vel, mode, layout, 0, c->cat.n);
        if (rc != ASTROZ_OK) return rc;
        return propagate_host_wait(c);
    }
    // one call, every device: each shard computes its satellite range and copies it over its own PCIe link into its
    // slice of the caller's block; no collective is needed for a host-resident result
    return for_each_shard(c, [&](size_t k) -> int32_t {
        Constellation *sh = c->shards[k];
        const int32_t q = propagate_host_queue(sh, jd, fr, n_times, pos, vel, mode, layout, c->shardRow0[k], c->cat.n);
        const int32_t w = propagate_host_wait(sh);   // also drains what was queued before a failure
        return q != ASTROZ_OK ? q : w;
    });
}

int32_t astroz_cuda_constellation_reset_carry(astroz_constellation_t h) { return h ? ASTROZ_OK : ASTROZ_NULL_POINTER; }

int32_t astroz_cuda_constellation_synchronize(astroz_constellation_t h) {
    if (!h) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    if (c->multi()) {
        for (Constellation *sh : c->shards) {
            const int32_t rc = astroz_cuda_constellation_synchronize(sh);
            if (rc != ASTROZ_OK) return rc;
        }
        return ASTROZ_OK;
    }
    if (!c->stream) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    AZ_CUDA(cudaStreamSynchronize(c->stream));
    AZ_CUDA(cudaStreamSynchronize(c->copyStream));
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_set_timing(astroz_constellation_t h, int32_t enabled) {
    if (!h) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    c->timing = enabled != 0;
    c->timed = false;
    for (Constellation *sh : c->shards) {
        sh->timing = c->timing;
        sh->timed = false;
    }
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_last_kernel_ms(astroz_constellation_t h, float ms[3]) {
    if (!h || !ms) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    ms[0] = ms[1] = ms[2] = 0.f;
    if (c->multi()) {  // the devices run side by side: the slowest one is the call's kernel time
        for (Constellation *sh : c->shards) {
            float one[3];
            const int32_t rc = astroz_cuda_constellation_last_kernel_ms(sh, one);
            if (rc != ASTROZ_OK) return rc;
            for (int k = 0; k < 3; ++k) ms[k] = std::max(ms[k], one[k]);
        }
        return ASTROZ_OK;
    }
    if (!c->timed) return ASTROZ_NOT_INITIALIZED;
    AZ_CUDA(cudaSetDevice(c->device));
    AZ_CUDA(cudaEventSynchronize(c->ev[1]));
    AZ_CUDA(cudaEventSynchronize(c->ev[3]));
    AZ_CUDA(cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
    AZ_CUDA(cudaEventElapsedTime(&ms[2], c->ev[2], c->ev[3]));
    if (c->spanTimed) {  // the two grids of a mixed call overlap: ms[1] is the span of the whole call
        AZ_CUDA(cudaEventSynchronize(c->ev[5]));
        AZ_CUDA(cudaEventElapsedTime(&ms[1], c->ev[4], c->ev[5]));
    } else {
        ms[1] = ms[0] + ms[2];
    }
    return ASTROZ_OK;
}

// ---- stateless near-earth path -----------------------------------------------------------------
static int32_t sgp4_into_common(Constellation *c, const double *times, uint32_t nt, const double *epoch_offsets,
                                double *dPos, double *dVel, int mode, double reference_jd, int layout, cudaStream_t s,
                                uint32_t recStride = 3, const uint8_t *mask = nullptr, uint32_t outNumSats = 0) {
    const uint32_t ns = c->cat.nSgp4;
    const uint32_t padded = c->cat.sgp4Padded();
    int32_t rc = reserve_time(c, nt);
    if (rc != ASTROZ_OK) return rc;
    const size_t cap = c->dTime.cap / 4;
    double *tb = c->hTime, *gs = c->hTime + 2 * (size_t)nt, *gc = c->hTime + 3 * (size_t)nt;
    for (uint32_t t = 0; t < nt; ++t) {
        tb[t] = times[t];
        if (mode != 0) {  // src/Constellation.zig:573-581
            const double gm = az::julian_to_gmst(reference_jd + times[t] / 1440.0);
            gs[t] = std::sin(gm);
            gc[t] = std::cos(gm);
        }
    }
    AZ_CUDA(cudaMemcpyAsync(c->dTime.p, tb, (size_t)nt * 8, cudaMemcpyHostToDevice, s));
    if (mode != 0) {
        AZ_CUDA(cudaMemcpyAsync(c->dTime.p + 2 * cap, gs, (size_t)nt * 8, cudaMemcpyHostToDevice, s));
        AZ_CUDA(cudaMemcpyAsync(c->dTime.p + 3 * cap, gc, (size_t)nt * 8, cudaMemcpyHostToDevice, s));
    }
    if (c->toffPending) {  // the previous call's upload of the offsets must have left the staging buffer
        AZ_CUDA(cudaEventSynchronize(c->toffCopied));
        c->toffPending = false;
    }
    if (padded > c->hToffCap) {
        if (c->hToffCall) cudaFreeHost(c->hToffCall);
        c->hToffCall = nullptr;
        AZ_CUDA(cudaMallocHost(&c->hToffCall, (size_t)padded * 8));
        c->hToffCap = padded;
    }
    for (uint32_t i = 0; i < padded; ++i) c->hToffCall[i] = epoch_offsets[std::min(i, ns - 1)];
    AZ_CUDA(c->dToffCall.reserve(padded));
    AZ_CUDA(cudaMemcpyAsync(c->dToffCall.p, c->hToffCall, (size_t)padded * 8, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaEventRecord(c->toffCopied, s));
    c->toffPending = true;
    AZ_CUDA(cudaEventRecord(c->timeCopied, s));
    c->timePending = true;

    az::GridArgs a;
    a.g = c->g;
    a.sgp4Tiles = c->dTiles.p;
    a.toff = c->dToffCall.p;
    a.orig = c->dIdentity.p;  // satellite i -> output row i (src/Constellation.zig:561-565)
    a.nSats = ns;
    a.tbase = c->dTime.p;
    a.gsin = c->dTime.p + 2 * cap;
    a.gcos = c->dTime.p + 3 * cap;
    a.nTimes = nt;
    a.pos = dPos;
    a.vel = dVel;
    a.outNumSats = outNumSats ? outNumSats : ns;
    a.recStride = recStride;
    if (mask) {  // pageable host bytes: the copy is staged by the driver before the call returns
        AZ_CUDA(c->dMask.reserve(ns));
        AZ_CUDA(cudaMemcpyAsync(c->dMask.p, mask, ns, cudaMemcpyHostToDevice, s));
        a.mask = c->dMask.p;
    }
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    AZ_CUDA(az::launch_sgp4_grid(a, mode, layout, s, c->variant));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
    c->timed = c->timing;
    c->spanTimed = false;
    return ASTROZ_OK;
}

int32_t astroz_cuda_sgp4_propagate_into_device(astroz_constellation_t h, const double *times, uint32_t n_times,
                                               const double *epoch_offsets, double *d_pos, double *d_vel, int32_t mode,
                                               double reference_jd, int32_t layout, const uint8_t *satellite_mask,
                                               uint32_t out_num_sats, void *stream) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    int32_t rc = check_args(c, times, epoch_offsets, d_pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    if (n_times == 0 || c->cat.nSgp4 == 0) return ASTROZ_OK;
    if (out_num_sats && out_num_sats < c->cat.nSgp4) {
        g_lastError = "out_num_sats smaller than the number of near-earth satellites";
        return ASTROZ_VALUE_ERROR;
    }
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return sgp4_into_common(c, times, n_times, epoch_offsets, d_pos, d_vel, mode, reference_jd, layout, s, 3,
                            satellite_mask, out_num_sats);
}

// Host-buffer form of the stateless near-earth path, in queue / wait halves like propagate_host_queue.  Near-earth
// satellite i of this handle lands in row rowOffset + i of a host block with `rows` rows; epoch_offsets / mask are
// already offset to this handle's first satellite.
static int32_t sgp4_into_host_queue(Constellation *c, const double *times, uint32_t n_times, const double *epoch_offsets,
                                    double *pos, double *vel, int32_t mode, double reference_jd, int32_t layout,
                                    const uint8_t *mask, uint32_t rows, uint32_t rowOffset) {
    const uint32_t ns = c->cat.nSgp4;
    if (n_times == 0 || ns == 0) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    // the near-earth rows are computed as a dense (ns, n_times) block on the device and placed in the caller's
    // (possibly wider) block with one strided copy per array; rows that belong to other satellites are never touched
    const size_t dense = (size_t)ns * n_times * 3;
    AZ_CUDA(c->dPos.reserve(dense));
    if (vel) AZ_CUDA(c->dVel.reserve(dense));
    const bool strided = (layout == 1 && rows != ns);
    auto host_at = [&](double *base) {
        return base + (layout == 0 ? (size_t)rowOffset * n_times * 3 : (size_t)rowOffset * 3);
    };
    if (mask) {  // masked rows must keep the caller's contents: stage the caller's rows, overwrite the active ones
        for (int which = 0; which < (vel ? 2 : 1); ++which) {
            double *dst = which ? c->dVel.p : c->dPos.p;
            const double *src = host_at(which ? vel : pos);
            if (!strided) AZ_CUDA(cudaMemcpyAsync(dst, src, dense * 8, cudaMemcpyHostToDevice, s));
            else AZ_CUDA(cudaMemcpy2DAsync(dst, (size_t)ns * 24, src, (size_t)rows * 24, (size_t)ns * 24, n_times,
                                           cudaMemcpyHostToDevice, s));
        }
    }
    int32_t rc = sgp4_into_common(c, times, n_times, epoch_offsets, c->dPos.p, vel ? c->dVel.p : nullptr, mode,
                                  reference_jd, layout, s, 3, mask, ns);
    if (rc != ASTROZ_OK) return rc;
    c->plan.clear();
    AZ_CUDA(cudaEventRecord(c->chunkDone[0], s));
    AZ_CUDA(cudaStreamWaitEvent(c->copyStream, c->chunkDone[0], 0));
    for (int which = 0; which < (vel ? 2 : 1); ++which) {
        double *dst = host_at(which ? vel : pos);
        const double *src = which ? c->dVel.p : c->dPos.p;
        const bool pg = is_pageable(which ? vel : pos);
        if (!strided) rc = deliver(c, pg, 0, src, dst, 1, dense * 8, dense * 8);
        else rc = deliver(c, pg, 0, src, dst, n_times, (size_t)ns * 24, (size_t)rows * 24);
        if (rc != ASTROZ_OK) return rc;
    }
    return ASTROZ_OK;
}

int32_t astroz_cuda_sgp4_propagate_into(astroz_constellation_t h, const double *times, uint32_t n_times,
                                        const double *epoch_offsets, double *pos, double *vel, int32_t mode,
                                        double reference_jd, int32_t layout, const uint8_t *satellite_mask,
                                        uint32_t out_num_sats) {
    Constellation *c = static_cast<Constellation *>(h);
    int32_t rc = check_args(c, times, epoch_offsets, pos, mode, layout);
    if (rc != ASTROZ_OK) return rc;
    const uint32_t ns = c->cat.nSgp4;
    if (n_times == 0 || ns == 0) return ASTROZ_OK;
    const uint32_t rows = out_num_sats ? out_num_sats : ns;
    if (rows < ns) {
        g_lastError = "out_num_sats smaller than the number of near-earth satellites";
        return ASTROZ_VALUE_ERROR;
    }
    if (!c->multi()) {
        rc = sgp4_into_host_queue(c, times, n_times, epoch_offsets, pos, vel, mode, reference_jd, layout, satellite_mask,
                                  rows, 0);
        if (rc != ASTROZ_OK) return rc;
        return propagate_host_wait(c);
    }
    return for_each_shard(c, [&](size_t k) -> int32_t {
        Constellation *sh = c->shards[k];
        const uint32_t near0 = c->shardNear0[k];
        const int32_t q = sgp4_into_host_queue(sh, times, n_times, epoch_offsets + near0, pos, vel, mode, reference_jd, layout,
                                               satellite_mask ? satellite_mask + near0 : nullptr, rows, near0);
        const int32_t w = propagate_host_wait(sh);
        return q != ASTROZ_OK ? q : w;
    });
}

int32_t astroz_cuda_sgp4_screen(astroz_constellation_t h, const double *times, uint32_t n_times,
                                const double *epoch_offsets, uint32_t target_idx, double threshold,
                                double reference_jd, double *out_min_dists, uint32_t *out_min_t) {
    (void)reference_jd;
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    if (!c || !times || !epoch_offsets || !out_min_dists || !out_min_t) return ASTROZ_NULL_POINTER;
    const uint32_t ns = c->cat.nSgp4;
    if (ns == 0 || n_times == 0) return ASTROZ_OK;
    if (target_idx >= ns) {
        g_lastError = "target index out of range";
        return ASTROZ_VALUE_ERROR;
    }
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t padded = c->cat.sgp4Padded();
    int32_t rc = reserve_time(c, n_times);
    if (rc != ASTROZ_OK) return rc;
    for (uint32_t t = 0; t < n_times; ++t) c->hTime[t] = times[t];
    AZ_CUDA(cudaMemcpyAsync(c->dTime.p, c->hTime, (size_t)n_times * 8, cudaMemcpyHostToDevice, s));
    if (c->toffPending) {  // the previous call's upload of the offsets must have left the staging buffer
        AZ_CUDA(cudaEventSynchronize(c->toffCopied));
        c->toffPending = false;
    }
    if (padded > c->hToffCap) {
        if (c->hToffCall) cudaFreeHost(c->hToffCall);
        c->hToffCall = nullptr;
        AZ_CUDA(cudaMallocHost(&c->hToffCall, (size_t)padded * 8));
        c->hToffCap = padded;
    }
    for (uint32_t i = 0; i < padded; ++i) c->hToffCall[i] = epoch_offsets[std::min(i, ns - 1)];
    AZ_CUDA(c->dToffCall.reserve(padded));
    AZ_CUDA(cudaMemcpyAsync(c->dToffCall.p, c->hToffCall, (size_t)padded * 8, cudaMemcpyHostToDevice, s));
    AZ_CUDA(cudaEventRecord(c->toffCopied, s));
    c->toffPending = true;
    AZ_CUDA(cudaEventRecord(c->timeCopied, s));
    c->timePending = true;
    // scratch: target track [nt][3] | minDist [ns] | minT [ns] (as doubles' worth of space)
    AZ_CUDA(c->dPos.reserve((size_t)n_times * 3 + 2 * (size_t)ns + 2));
    az::ScreenArgs a;
    a.sgp4Tiles = c->dTiles.p;
    a.toff = c->dToffCall.p;
    a.tbase = c->dTime.p;
    a.nSats = ns;
    a.nTimes = n_times;
    a.targetIdx = target_idx;
    a.thresholdSq = threshold * threshold;
    a.track = c->dPos.p;
    a.minDist = c->dPos.p + (size_t)n_times * 3;
    a.minT = reinterpret_cast<uint32_t *>(a.minDist + ns);
    a.g = c->g;
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], s));
    AZ_CUDA(az::launch_sgp4_screen(a, s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], s));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], s));
    c->timed = c->timing;
    c->spanTimed = false;
    AZ_CUDA(cudaMemcpyAsync(out_min_dists, a.minDist, (size_t)ns * 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaMemcpyAsync(out_min_t, a.minT, (size_t)ns * 4, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    return ASTROZ_OK;
}

// ---- all-vs-all coarse screen ----------------------------------------------------------------------
static int32_t coarse_screen_run(Constellation *c, const double *dPositions, uint32_t ns, uint32_t nt, int layout,
                                 double threshold, const uint8_t *dMask, uint32_t *dPairs, uint32_t *dT,
                                 uint32_t maxResults, uint64_t *count, cudaStream_t s) {
    constexpr uint32_t kBits = 16, kBatch = 128;  // 128 epochs per pass: 32 MB of bucket heads
    const uint32_t batch = std::min(kBatch, nt);
    AZ_CUDA(c->dHead.reserve((size_t)batch << kBits));
    AZ_CUDA(c->dNext.reserve((size_t)batch * ns));
    AZ_CUDA(c->dCount.reserve(1));
    AZ_CUDA(cudaMemsetAsync(c->dCount.p, 0, 8, s));
    for (uint32_t t0 = 0; t0 < nt; t0 += batch) {
        az::CoarseArgs a;
        a.pos = dPositions;
        a.validMask = dMask;
        a.nSats = ns;
        a.nTimes = nt;
        a.layout = layout;
        a.threshold = threshold;
        a.t0 = t0;
        a.tCount = std::min(batch, nt - t0);
        a.tableBits = kBits;
        a.head = c->dHead.p;
        a.next = c->dNext.p;
        a.pairs = dPairs;
        a.tIdx = dT;
        a.maxResults = maxResults;
        a.count = c->dCount.p;
        AZ_CUDA(az::launch_coarse_screen(a, s));
    }
    unsigned long long found = 0;
    AZ_CUDA(cudaMemcpyAsync(&found, c->dCount.p, 8, cudaMemcpyDeviceToHost, s));
    AZ_CUDA(cudaStreamSynchronize(s));
    *count = found;
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_coarse_screen_device(astroz_constellation_t h, const double *d_positions,
                                                       uint32_t num_sats, uint32_t num_times, int32_t layout,
                                                       double threshold, const uint8_t *d_valid_mask, uint32_t *d_pairs,
                                                       uint32_t *d_t_indices, uint32_t max_results, uint64_t *count) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    if (!c || !d_positions || !count || (max_results && (!d_pairs || !d_t_indices))) return ASTROZ_NULL_POINTER;
    if (layout < 0 || layout > 1 || !(threshold > 0.0)) {
        g_lastError = "coarse screen: layout must be 0/1 and threshold positive";
        return ASTROZ_VALUE_ERROR;
    }
    *count = 0;
    if (num_sats == 0 || num_times == 0) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    return coarse_screen_run(c, d_positions, num_sats, num_times, layout, threshold, d_valid_mask, d_pairs, d_t_indices,
                             max_results, count, c->stream);
}

int32_t astroz_cuda_sgp4_screen_all(astroz_constellation_t h, const double *times, uint32_t n_times,
                                    const double *epoch_offsets, double threshold, uint32_t *pairs, uint32_t *t_indices,
                                    uint32_t max_results, uint64_t *count) {
    Constellation *c = static_cast<Constellation *>(h);
    AZ_SINGLE(c);
    if (!c || !times || !epoch_offsets || !count || (max_results && (!pairs || !t_indices))) return ASTROZ_NULL_POINTER;
    if (!(threshold > 0.0)) return ASTROZ_VALUE_ERROR;
    *count = 0;
    const uint32_t ns = c->cat.nSgp4;
    if (ns == 0 || n_times == 0) return ASTROZ_OK;
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    AZ_CUDA(c->dPos.reserve((size_t)ns * n_times * 3));
    // positions only, time-major: a warp of the screen kernels then reads 32 neighbouring satellites of one epoch
    int32_t rc = sgp4_into_common(c, times, n_times, epoch_offsets, c->dPos.p, nullptr, ASTROZ_MODE_TEME, 0.0,
                                  ASTROZ_LAYOUT_TIME_MAJOR, s);
    if (rc != ASTROZ_OK) return rc;
    AZ_CUDA(c->dPairs.reserve((size_t)std::max<uint32_t>(max_results, 1) * 2));
    AZ_CUDA(c->dTIdx.reserve(std::max<uint32_t>(max_results, 1)));
    rc = coarse_screen_run(c, c->dPos.p, ns, n_times, ASTROZ_LAYOUT_TIME_MAJOR, threshold, nullptr, c->dPairs.p,
                           c->dTIdx.p, max_results, count, s);
    if (rc != ASTROZ_OK) return rc;
    const size_t stored = (size_t)std::min<uint64_t>(*count, max_results);
    if (stored) {
        AZ_CUDA(cudaMemcpyAsync(pairs, c->dPairs.p, stored * 8, cudaMemcpyDeviceToHost, s));
        AZ_CUDA(cudaMemcpyAsync(t_indices, c->dTIdx.p, stored * 4, cudaMemcpyDeviceToHost, s));
        AZ_CUDA(cudaStreamSynchronize(s));
    }
    return ASTROZ_OK;
}

// ---- single satellite ----------------------------------------------------------------------------
int32_t astroz_cuda_sgp4_init(const char *line1, const char *line2, int32_t grav, int32_t device, astroz_sgp4_t *out) {
    if (!line1 || !line2 || !out) return ASTROZ_NULL_POINTER;
    *out = nullptr;
    // python-sgp4 style code builds thousands of Satrec objects only to hand them to a SatrecArray: parsing and
    // classification happen here, on the host; device resources come with the first propagation of THIS satellite.
    {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            g_lastError = "no CUDA device available (this library has no CPU propagation path)";
            return ASTROZ_NO_DEVICE;
        }
        if (device < 0 || device >= count) {
            g_lastError = "device index out of range";
            return ASTROZ_VALUE_ERROR;
        }
    }
    Constellation *cc = new (std::nothrow) Constellation();
    if (!cc) return ASTROZ_ALLOC_FAILED;
    const char *l1[1] = {line1}, *l2[1] = {line2};
    const int brc = az::build_catalog(l1, l2, 1, grav, cc->cat);
    if (brc != az::kOk) {
        delete cc;
        return status_to_code(brc);
    }
    Sgp4Single *s = new (std::nothrow) Sgp4Single();
    if (!s) {
        delete cc;
        return ASTROZ_ALLOC_FAILED;
    }
    s->c = cc;
    s->device = device;
    s->epochJd = s->c->cat.epochs[0];
    s->deep = s->c->cat.nSdp4 == 1;
    {  // mean elements for the python-sgp4 attribute getters (bindings/python/src/satrec.zig:395-470)
        az::TleRecord t;
        az::NearEarth ne;
        double period = 0.0, perigee = 0.0;
        if (az::parse_tle(line1, line2, t) == az::kOk &&
            az::build_common(t, az::gravity(grav), ne, period, perigee) == az::kOk) {
            const double e[10] = {ne.ecco, ne.inclo, ne.nodeo, ne.argpo, ne.mo, ne.no_kozai, ne.bstar, ne.a, ne.no, ne.epochJd};
            std::memcpy(s->elements, e, sizeof e);
        }
    }
    *out = s;
    return ASTROZ_OK;
}

static int32_t ensure_open(Sgp4Single *s) {
    if (s->opened) return ASTROZ_OK;
    const int32_t rc = finish_create(s->c, s->device);
    if (rc == ASTROZ_OK) s->opened = true;
    return rc;
}

int32_t astroz_cuda_sgp4_elements(astroz_sgp4_t h, double *out10) {
    if (!h || !out10) return ASTROZ_NULL_POINTER;
    std::memcpy(out10, static_cast<Sgp4Single *>(h)->elements, sizeof(double) * 10);
    return ASTROZ_OK;
}

void astroz_cuda_sgp4_free(astroz_sgp4_t h) {
    Sgp4Single *s = static_cast<Sgp4Single *>(h);
    if (!s) return;
    delete s->c;
    delete s;
}

int32_t astroz_cuda_sgp4_is_deep_space(astroz_sgp4_t h) { return h && static_cast<Sgp4Single *>(h)->deep ? 1 : 0; }

int32_t astroz_cuda_sgp4_epoch(astroz_sgp4_t h, double *epoch_jd) {
    if (!h || !epoch_jd) return ASTROZ_NULL_POINTER;
    *epoch_jd = static_cast<Sgp4Single *>(h)->epochJd;
    return ASTROZ_OK;
}

int32_t astroz_cuda_sgp4_propagate_batch(astroz_sgp4_t h, const double *times, double *results, uint32_t count) {
    Sgp4Single *s = static_cast<Sgp4Single *>(h);
    if (!s || !times || !results) return ASTROZ_NULL_POINTER;
    if (count == 0) return ASTROZ_OK;
    {
        const int32_t orc = ensure_open(s);
        if (orc != ASTROZ_OK) return orc;
    }
    Constellation *c = s->c;
    AZ_CUDA(cudaSetDevice(c->device));
    const bool fast = !s->deep && count >= 64;
    std::vector<double> pos(fast ? 0 : (size_t)count * 3), vel(fast ? 0 : (size_t)count * 3);
    int32_t rc;
    if (!s->deep) {
        // near earth: the time-parallel kernel writes x y z vx vy vz records straight into one block that
        // is copied to `results` in a single transfer (src/c_api/sgp4.zig:60-100 layout)
        const double zero = 0.0;
        if (count >= 64) {
            AZ_CUDA(c->dPos.reserve((size_t)count * 6));
            rc = sgp4_into_common(c, times, count, &zero, c->dPos.p, c->dPos.p + 3, ASTROZ_MODE_TEME, 0.0,
                                  ASTROZ_LAYOUT_SATELLITE_MAJOR, c->stream, 6);
            if (rc != ASTROZ_OK) return rc;
            AZ_CUDA(cudaMemcpyAsync(results, c->dPos.p, (size_t)count * 48, cudaMemcpyDeviceToHost, c->stream));
            AZ_CUDA(cudaStreamSynchronize(c->stream));
            return ASTROZ_OK;
        }
        rc = astroz_cuda_sgp4_propagate_into(c, times, count, &zero, pos.data(), vel.data(), ASTROZ_MODE_TEME, 0.0,
                                             ASTROZ_LAYOUT_SATELLITE_MAJOR, nullptr, 0);
        if (rc != ASTROZ_OK) return rc;
    } else {
        // deep space: minutes since epoch go to the kernel directly (no Julian-date round trip)
        int32_t rt = reserve_time(c, count);
        if (rt != ASTROZ_OK) return rt;
        double reach = 0.0;
        for (uint32_t i = 0; i < count; ++i) {
            c->hTime[i] = times[i];
            reach = std::max(reach, std::fabs(times[i]));
        }
        cudaStream_t st = c->stream;
        AZ_CUDA(cudaMemcpyAsync(c->dTime.p, c->hTime, (size_t)count * 8, cudaMemcpyHostToDevice, st));
        AZ_CUDA(cudaEventRecord(c->timeCopied, st));
        c->timePending = true;
        rc = ensure_lattice(c, (int)std::floor(reach / az::kStepp) + 2, st);
        if (rc != ASTROZ_OK) return rc;
        const size_t total = (size_t)count * 3;
        DevBuf<uint8_t> dSt;
        AZ_CUDA(c->dPos.reserve(total));
        AZ_CUDA(c->dVel.reserve(total));
        AZ_CUDA(dSt.reserve(count));
        az::GridArgs a;
        a.g = c->g;
        a.sdp4 = c->dSdp4.p;
        a.orig = c->dSdp4Orig.p;
        a.nSats = 1;
        a.lattice = c->dLattice.p;
        a.latticeNodes = c->latticeNodes;
        a.tsince = c->dTime.p;
        a.nTimes = count;
        a.pos = c->dPos.p;
        a.vel = c->dVel.p;
        a.status = dSt.p;
        a.outNumSats = 1;
        std::vector<uint8_t> cell(count);
        cudaError_t e = az::launch_sdp4_grid(a, ASTROZ_MODE_TEME, ASTROZ_LAYOUT_SATELLITE_MAJOR, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pos.data(), c->dPos.p, total * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(vel.data(), c->dVel.p, total * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(cell.data(), dSt.p, count, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        dSt.release();
        if (e != cudaSuccess) return cuda_fail(e, "single-satellite deep-space propagate");
        rc = ASTROZ_OK;
        for (uint32_t i = 0; i < count; ++i)
            if (cell[i] != 0) rc = status_to_code(cell[i]);  // failing cells stay zero-filled; last failure reported
    }
    for (uint32_t i = 0; i < count; ++i) {
        double *r = results + (size_t)i * 6;
        r[0] = pos[i * 3]; r[1] = pos[i * 3 + 1]; r[2] = pos[i * 3 + 2];
        r[3] = vel[i * 3]; r[4] = vel[i * 3 + 1]; r[5] = vel[i * 3 + 2];
    }
    return rc;
}

int32_t astroz_cuda_sgp4_array(astroz_sgp4_t h, const double *jd, const double *fr, double epoch_jd, double *results,
                               uint32_t count) {
    Sgp4Single *s = static_cast<Sgp4Single *>(h);
    if (!s || !jd || !fr || !results) return ASTROZ_NULL_POINTER;
    if (count == 0) return ASTROZ_OK;
    {
        const int32_t orc = ensure_open(s);
        if (orc != ASTROZ_OK) return orc;
    }
    Constellation *c = s->c;
    if (s->deep || count < 64) {  // deep space / tiny: host-side tsince, then the batch entry point
        std::vector<double> ts(count);
        for (uint32_t i = 0; i < count; ++i) ts[i] = ((jd[i] + fr[i]) - epoch_jd) * 1440.0;
        return astroz_cuda_sgp4_propagate_batch(h, ts.data(), results, count);
    }
    AZ_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->cacheValid = false;
    // A long axis ("1 year at one second" = 31.5 M epochs: 0.5 GB of jd/fr in, 1.5 GB of records out) is cut into
    // chunks on a two-slot pipeline: while chunk k is propagated, chunk k+1's epochs are staged and uploaded and chunk
    // k-1's records travel back, so the call runs at the PCIe rate of its 48 B/epoch result instead of the sum of three
    // serial phases.  jd / fr usually are pageable (numpy): the copy pool stages them into pinned slots.
    constexpr uint32_t kChunk = 1u << 21;   // epochs per chunk: 32 MB of epochs, 96 MB of records
    const uint32_t nChunks = (count + kChunk - 1) / kChunk;
    const uint32_t chunk = ((count + nChunks - 1) / nChunks + 31) / 32 * 32;   // balanced: no short last chunk
    const int slots = nChunks > 1 ? 2 : 1;
    AZ_CUDA(c->dTime.reserve((size_t)chunk * 2 * slots));
    AZ_CUDA(c->dPos.reserve((size_t)chunk * 6 * slots));
    const bool inPageable = is_pageable(jd) || is_pageable(fr);
    const bool outPageable = is_pageable(results);
    if ((inPageable || outPageable) && !c->ring) {
        AZ_CUDA(cudaMallocHost(&c->ring, kPieceBytes * kRingSlots));
        for (auto &e : c->ringEv) AZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // pinned staging inside the ring allocation (96 MB): two 16 MB input slots [jd | fr]; outputs of a pageable caller
    // are not staged here (a 96 MB chunk does not fit) -- they go through cudaMemcpyAsync's own staging
    CopyPool &pool = CopyPool::get();
    cudaEvent_t kdone[2] = {c->chunkDone[0], c->chunkDone[1]}, ddone[2] = {c->chunkDone[2], c->chunkDone[3]};
    const uint32_t inChunk = inPageable ? std::min<uint32_t>(chunk, (uint32_t)(kPieceBytes / 16)) : chunk;  // staging granule
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[0], st));
    uint32_t granule = 0;
    for (uint32_t k = 0; k < nChunks; ++k) {
        const int slot = (int)(k & 1u) % slots;
        const uint32_t t0 = k * chunk, n = std::min(chunk, count - t0);
        double *dJd = c->dTime.p + (size_t)slot * chunk * 2, *dFr = dJd + chunk;
        double *dOut = c->dPos.p + (size_t)slot * chunk * 6;
        if (k >= (uint32_t)slots) AZ_CUDA(cudaEventSynchronize(ddone[slot]));  // chunk k-2 has left this slot
        if (inPageable) {   // stage through pinned memory in granules of the ring slot size, copy pool does the memcpy
            for (uint32_t g0 = 0; g0 < n; g0 += inChunk, ++granule) {
                const uint32_t gn = std::min(inChunk, n - g0);
                const int ri = (int)(granule % kRingSlots);
                char *stage = c->ring + (size_t)ri * kPieceBytes;
                if (granule >= (uint32_t)kRingSlots) AZ_CUDA(cudaEventSynchronize(c->ringEv[ri]));  // its last upload is done
                pool.copy(stage, reinterpret_cast<const char *>(jd + t0 + g0), 1, (size_t)gn * 8, (size_t)gn * 8);
                pool.copy(stage + (size_t)gn * 8, reinterpret_cast<const char *>(fr + t0 + g0), 1, (size_t)gn * 8, (size_t)gn * 8);
                AZ_CUDA(cudaMemcpyAsync(dJd + g0, stage, (size_t)gn * 8, cudaMemcpyHostToDevice, st));
                AZ_CUDA(cudaMemcpyAsync(dFr + g0, stage + (size_t)gn * 8, (size_t)gn * 8, cudaMemcpyHostToDevice, st));
                AZ_CUDA(cudaEventRecord(c->ringEv[ri], st));
            }
        } else {
            AZ_CUDA(cudaMemcpyAsync(dJd, jd + t0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            AZ_CUDA(cudaMemcpyAsync(dFr, fr + t0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        }
        az::GridArgs a;
        a.g = c->g;
        a.sgp4Tiles = c->dTiles.p;
        a.toff = c->dToff.p;
        a.orig = c->dIdentity.p;
        a.nSats = 1;
        a.jdArr = dJd;
        a.frArr = dFr;
        a.tbase = dJd;
        a.epochJd = epoch_jd;
        a.nTimes = n;
        a.pos = dOut;
        a.vel = dOut + 3;
        a.outNumSats = 1;
        a.recStride = 6;
        AZ_CUDA(az::launch_sgp4_grid(a, ASTROZ_MODE_TEME, ASTROZ_LAYOUT_SATELLITE_MAJOR, st, c->variant));
        AZ_CUDA(cudaEventRecord(kdone[slot], st));
        AZ_CUDA(cudaStreamWaitEvent(c->copyStream, kdone[slot], 0));
        AZ_CUDA(cudaMemcpyAsync(results + (size_t)t0 * 6, dOut, (size_t)n * 48, cudaMemcpyDeviceToHost, c->copyStream));
        AZ_CUDA(cudaEventRecord(ddone[slot], c->copyStream));
    }
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[1], st));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[2], st));
    if (c->timing) AZ_CUDA(cudaEventRecord(c->ev[3], st));
    c->timed = c->timing;
    c->spanTimed = false;
    AZ_CUDA(cudaStreamSynchronize(c->copyStream));
    AZ_CUDA(cudaStreamSynchronize(st));
    return ASTROZ_OK;
}

int32_t astroz_cuda_sgp4_propagate(astroz_sgp4_t h, double tsince, double pos[3], double vel[3]) {
    if (!h || !pos || !vel) return ASTROZ_NULL_POINTER;
    double r[6];
    int32_t rc = astroz_cuda_sgp4_propagate_batch(h, &tsince, r, 1);
    std::memcpy(pos, r, 24);
    std::memcpy(vel, r + 3, 24);
    return rc;
}

// ---- result blocks placed next to the GPUs that fill them ---------------------------------------------------------
// A multi-device handle writes one host block from several GPUs at once.  A plain pinned allocation sits on the NUMA
// node of the thread that made it, so half the GPUs of a two-socket box write across the socket interconnect and all of
// them into one node's memory controllers (measured: 93 GB/s for 8 GPUs against 315 GB/s when every GPU writes node-local
// memory).  astroz_cuda_constellation_host_block maps the block anonymously, binds each device's slice of a
// satellite-major block to the NUMA node of that device (mbind), touches it, and page-locks the whole range.
static int device_numa_node(int device) {
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1;
    }
    for (char *p = bus; *p; ++p) *p = (char)std::tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}

static void bind_to_node(char *begin, char *end, int node) {
#ifdef SYS_mbind
    if (node < 0 || node >= 64) return;
    const long page = sysconf(_SC_PAGESIZE);
    char *b = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(begin) + page - 1) / page * page);
    char *e = reinterpret_cast<char *>(reinterpret_cast<uintptr_t>(end) / page * page);
    if (e <= b) return;
    unsigned long mask = 1ul << node;
    (void)syscall(SYS_mbind, b, (unsigned long)(e - b), 2 /* MPOL_BIND */, &mask, 64ul, 0u);  // best effort
#else
    (void)begin; (void)end; (void)node;
#endif
}

int32_t astroz_cuda_constellation_host_block(astroz_constellation_t h, uint32_t n_times, int32_t layout, double **out) {
    if (!h || !out) return ASTROZ_NULL_POINTER;
    *out = nullptr;
    Constellation *c = static_cast<Constellation *>(h);
    const size_t bytes = std::max<size_t>((size_t)c->cat.n * n_times * 24, 4096);
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return ASTROZ_ALLOC_FAILED;
    char *base = static_cast<char *>(p);
    if (c->multi() && layout == ASTROZ_LAYOUT_SATELLITE_MAJOR) {
        for (size_t k = 0; k < c->shards.size(); ++k)
            bind_to_node(base + (size_t)c->shardRow0[k] * n_times * 24, base + (size_t)c->shardRow0[k + 1] * n_times * 24,
                         device_numa_node(c->shards[k]->device));
    } else if (!c->multi()) {
        bind_to_node(base, base + bytes, device_numa_node(c->device));
    }   // time-major over several devices: rows interleave, the default (first-touch) placement stays
    {   // touch every page so the placement happens now, on the bound node, not at the first DMA
        const long page = sysconf(_SC_PAGESIZE);
        for (size_t o = 0; o < bytes; o += (size_t)page) base[o] = 0;
    }
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        munmap(p, bytes);
        return cuda_fail(e, "cudaHostRegister");
    }
    {
        std::lock_guard<std::mutex> g(g_blockMutex);
        g_blocks[p] = bytes;
    }
    *out = static_cast<double *>(p);
    return ASTROZ_OK;
}

// ---- caller-owned buffers: explicit page-locking ------------------------------------------------------------
int32_t astroz_cuda_host_register(void *p, size_t bytes) {
    if (!p) return ASTROZ_NULL_POINTER;
    AZ_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return ASTROZ_OK;
}
int32_t astroz_cuda_host_unregister(void *p) {
    if (!p) return ASTROZ_NULL_POINTER;
    AZ_CUDA(cudaHostUnregister(p));
    return ASTROZ_OK;
}

// ---- multi-device handles --------------------------------------------------------------------------------
int32_t astroz_cuda_constellation_devices(astroz_constellation_t h, int32_t *n_devices, int32_t *device_ids,
                                          uint32_t *first_rows) {
    if (!h || !n_devices) return ASTROZ_NULL_POINTER;
    Constellation *c = static_cast<Constellation *>(h);
    if (!c->multi()) {
        *n_devices = 1;
        if (device_ids) device_ids[0] = c->device;
        if (first_rows) { first_rows[0] = 0; first_rows[1] = c->cat.n; }
        return ASTROZ_OK;
    }
    *n_devices = (int32_t)c->shards.size();
    for (size_t k = 0; k < c->shards.size(); ++k) {
        if (device_ids) device_ids[k] = c->shards[k]->device;
        if (first_rows) first_rows[k] = c->shardRow0[k];
    }
    if (first_rows) first_rows[c->shards.size()] = c->cat.n;
    return ASTROZ_OK;
}

// Peer mappings between every pair of distinct devices of the handle (cudaMalloc memory of one is then directly
// addressable from kernels on the other: NVLink 5 loads/stores).  Idempotent.
static int32_t enable_peers(Constellation *c) {
    for (Constellation *a : c->shards)
        for (Constellation *b : c->shards) {
            if (a->device == b->device) continue;
            int can = 0;
            AZ_CUDA(cudaDeviceCanAccessPeer(&can, a->device, b->device));
            if (!can) {
                g_lastError = "devices " + std::to_string(a->device) + " and " + std::to_string(b->device) +
                              " have no peer access (no NVLink / P2P path)";
                return ASTROZ_CUDA_ERROR;
            }
            AZ_CUDA(cudaSetDevice(a->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
        }
    return ASTROZ_OK;
}

int32_t astroz_cuda_constellation_propagate_replicated(astroz_constellation_t h, const double *jd, const double *fr,
                                                       uint32_t n_times, int32_t velocities, double **d_pos,
                                                       double **d_vel) {
    Constellation *c = static_cast<Constellation *>(h);
    if (!c || !jd || !fr || !d_pos || (velocities && !d_vel)) return ASTROZ_NULL_POINTER;
    const uint32_t n = c->cat.n;
    const size_t total = (size_t)n * n_times * 3;
    if (!c->multi()) {  // one device: the block is simply left in HBM
        if (n_times == 0 || n == 0) return ASTROZ_OK;
        AZ_CUDA(cudaSetDevice(c->device));
        AZ_CUDA(c->dFullPos.reserve(total));
        if (velocities) AZ_CUDA(c->dFullVel.reserve(total));
        int32_t rc = astroz_cuda_constellation_propagate_device(c, jd, fr, n_times, c->dFullPos.p,
                                                                velocities ? c->dFullVel.p : nullptr, nullptr,
                                                                ASTROZ_MODE_TEME, ASTROZ_LAYOUT_SATELLITE_MAJOR, n, 0, nullptr);
        if (rc != ASTROZ_OK) return rc;
        AZ_CUDA(cudaStreamSynchronize(c->stream));
        d_pos[0] = c->dFullPos.p;
        if (velocities) d_vel[0] = c->dFullVel.p;
        return ASTROZ_OK;
    }
    if (n_times == 0 || n == 0) return ASTROZ_OK;
    int32_t rc = enable_peers(c);
    if (rc != ASTROZ_OK) return rc;
    const size_t ns = c->shards.size();
    if (ns > (size_t)az::kMaxPeers) {
        g_lastError = "at most 8 devices (one NVSwitch domain)";
        return ASTROZ_VALUE_ERROR;
    }
    for (Constellation *sh : c->shards) {  // every device holds the whole block
        AZ_CUDA(cudaSetDevice(sh->device));
        AZ_CUDA(sh->dFullPos.reserve(total));
        if (velocities) AZ_CUDA(sh->dFullVel.reserve(total));
    }
    void *pp[az::kMaxPeers] = {}, *pv[az::kMaxPeers] = {};
    for (size_t k = 0; k < ns; ++k) {
        pp[k] = c->shards[k]->dFullPos.p;
        pv[k] = velocities ? c->shards[k]->dFullVel.p : nullptr;
    }
    // one fused launch per device: its rows are stored, run by run, into every device's copy of the block
    int32_t first = ASTROZ_OK;
    for (size_t k = 0; k < ns; ++k) {
        rc = astroz_cuda_constellation_propagate_gather(c->shards[k], jd, fr, n_times, pp, velocities ? pv : nullptr,
                                                        (uint32_t)ns, nullptr, nullptr, n, c->shardRow0[k], nullptr);
        if (rc != ASTROZ_OK && first == ASTROZ_OK) first = rc;
    }
    for (size_t k = 0; k < ns; ++k) {  // all stores have landed once every device's stream has drained
        rc = propagate_host_wait(c->shards[k]);
        if (rc != ASTROZ_OK && first == ASTROZ_OK) first = rc;
        d_pos[k] = c->shards[k]->dFullPos.p;
        if (velocities) d_vel[k] = c->shards[k]->dFullVel.p;
    }
    return first;
}

int32_t astroz_cuda_fp64_peak(int32_t device, double *tflops) {
    if (!tflops) return ASTROZ_NULL_POINTER;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_lastError = "no CUDA device available";
        return ASTROZ_NO_DEVICE;
    }
    AZ_CUDA(cudaSetDevice(device));
    double flops = 0;
    AZ_CUDA(az::measure_fp64_peak(&flops));
    *tflops = flops * 1e-12;
    return ASTROZ_OK;
}

int32_t astroz_cuda_fp64_pipe_peak(int32_t device, double *tflops) {
    if (!tflops) return ASTROZ_NULL_POINTER;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        g_lastError = "no CUDA device available";
        return ASTROZ_NO_DEVICE;
    }
    AZ_CUDA(cudaSetDevice(device));
    double flops = 0;
    AZ_CUDA(az::fp64_pipe_peak(&flops));
    *tflops = flops * 1e-12;
    return ASTROZ_OK;
}

#pragma GCC visibility pop
}  // extern "C"
