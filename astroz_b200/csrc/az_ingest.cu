// az_ingest.cu -- device-side element initialisation (K5): mean elements -> propagation constants, in HBM.
//
// Replaces, for element sets that are already device resident (Monte-Carlo draws, OMM streams), the host loop of
// src/Constellation.zig:101-200: Sgp4.initElements (src/Sgp4.zig:108-417) per satellite, Sdp4.initElements
// (src/Sdp4.zig:174-657: gstime, dscom, dsinit) for the ones whose period exceeds 225 min, classification into the
// near-earth tile table and the deep-space record list in catalog order, and the 8-wide padding of the last tile
// (src/Constellation.zig:146).  The arithmetic is the same source as the host builders (az_elements.hpp compiled for
// the device); this file only adds the classification scan and the table scatter.
//
//   K5a classify : one thread per element set: run the common init, flag near (0) / deep (1) / failed, count the
//                  flags per 256-thread block, record the first failing catalog index.
//   K5b offsets  : one block: exclusive scan of the per-block counts -> per-block output bases, totals.
//   K5c build    : one thread per element set: rank inside the block by warp ballots, full init, scatter into the
//                  tile columns / record list; the owner of the last near-earth satellite also fills the padding.
#include <cuda_runtime.h>
#include <stdint.h>

#include "az_ingest.cuh"

namespace az {

namespace {

constexpr int kIngestThreads = 256;

__device__ __forceinline__ TleRecord load_elements(const IngestArgs &a, uint32_t i) {
    TleRecord t;
    t.satnum = i;
    t.epochJd = a.epochJd[i];
    t.revPerDay = a.revPerDay[i];
    t.ecc = a.ecc[i];
    t.inclDeg = a.inclDeg[i];
    t.raanDeg = a.raanDeg[i];
    t.argpDeg = a.argpDeg[i];
    t.maDeg = a.maDeg[i];
    t.bstar = a.bstar[i];
    return t;
}

// flag: 0 near earth, 1 deep space, 2 failed
__global__ void __launch_bounds__(kIngestThreads) ingest_classify_kernel(IngestArgs a) {
    const uint32_t i = blockIdx.x * kIngestThreads + threadIdx.x;
    int flag = 3;  // out of range
    if (i < a.n) {
        const TleRecord t = load_elements(a, i);
        NearEarth ne;
        double period = 0.0, perigee = 0.0;
        const int rc = build_common(t, a.grav, ne, period, perigee);
        if (rc != kOk) {
            flag = 2;
            // first failure in catalog order wins, like the host loop (src/Constellation.zig:115-126)
            atomicMin(a.firstFail, ((unsigned long long)i << 8) | (unsigned)rc);
        } else {
            flag = period > 225.0 ? 1 : 0;
        }
        a.flags[i] = (uint8_t)flag;
    }
    const int nNear = __syncthreads_count(flag == 0);
    const int nDeep = __syncthreads_count(flag == 1);
    if (threadIdx.x == 0) {
        a.blockNear[blockIdx.x] = (uint32_t)nNear;
        a.blockDeep[blockIdx.x] = (uint32_t)nDeep;
    }
}

// single block: exclusive scan of both per-block count arrays, in place; totals -> a.totals[0..1]
__global__ void __launch_bounds__(1024) ingest_offsets_kernel(IngestArgs a, uint32_t nBlocks) {
    __shared__ uint32_t warpSum[2][32];
    __shared__ uint32_t carry[2];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v[2] = {i < nBlocks ? a.blockNear[i] : 0u, i < nBlocks ? a.blockDeep[i] : 0u};
        uint32_t incl[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint32_t x = v[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            incl[k] = x;
            if (lane == 31) warpSum[k][warp] = x;
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                uint32_t x = warpSum[k][lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                    if (lane >= (uint32_t)d) x += y;
                }
                warpSum[k][lane] = x;  // inclusive over warps
            }
        }
        __syncthreads();
        uint32_t excl[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
            excl[k] = carry[k] + (warp ? warpSum[k][warp - 1] : 0u) + incl[k] - v[k];
        if (i < nBlocks) {
            a.blockNear[i] = excl[0];
            a.blockDeep[i] = excl[1];
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry[0] = excl[0] + v[0];
            carry[1] = excl[1] + v[1];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.totals[0] = carry[0];
        a.totals[1] = carry[1];
    }
}

__device__ __forceinline__ void write_tile_lane(double *tiles, uint32_t slot, const double *cols) {
    double *tile = tiles + (size_t)(slot / kTileSats) * kSgp4TileDoubles + (slot % kTileSats);
#pragma unroll
    for (int c = 0; c < kSgp4Cols; ++c) tile[c * kTileSats] = cols[c];
}

__global__ void __launch_bounds__(kIngestThreads) ingest_build_kernel(IngestArgs a) {
    __shared__ uint32_t warpNear[kIngestThreads / 32], warpDeep[kIngestThreads / 32];
    const uint32_t i = blockIdx.x * kIngestThreads + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int flag = i < a.n ? a.flags[i] : 3;
    const uint32_t bNear = __ballot_sync(0xffffffffu, flag == 0), bDeep = __ballot_sync(0xffffffffu, flag == 1);
    if (lane == 0) {
        warpNear[warp] = __popc(bNear);
        warpDeep[warp] = __popc(bDeep);
    }
    __syncthreads();
    uint32_t rankNear = a.blockNear[blockIdx.x], rankDeep = a.blockDeep[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) {
        rankNear += warpNear[w];
        rankDeep += warpDeep[w];
    }
    const uint32_t below = (1u << lane) - 1u;
    rankNear += __popc(bNear & below);
    rankDeep += __popc(bDeep & below);
    if (flag > 1) {
        if (i < a.n) a.classes[i] = -1;
        return;
    }
    const TleRecord t = load_elements(a, i);
    if (flag == 0) {
        NearEarth ne;
        build_near_earth(t, a.grav, ne);
        double cols[kSgp4Cols];
        sgp4_columns(ne, cols);
        write_tile_lane(a.tiles, rankNear, cols);
        a.sgp4Orig[rankNear] = i;
        a.identity[rankNear] = rankNear;
        a.classes[i] = 0;
        const uint32_t nNear = a.totals[0];
        if (rankNear == nNear - 1) {  // padding lanes of the last tile repeat this satellite (src/Constellation.zig:146)
            const uint32_t padded = (nNear + kTileSats - 1) / kTileSats * kTileSats;
            for (uint32_t s = nNear; s < padded; ++s) {
                write_tile_lane(a.tiles, s, cols);
                a.sgp4Orig[s] = i;
                a.identity[s] = rankNear;
            }
        }
    } else {
        DeepSpace ds;
        const int rc = build_deep_space(t, a.grav, ds);
        if (rc != kOk) {  // cannot happen after build_common passed, kept for symmetry with the host loop
            atomicMin(a.firstFail, ((unsigned long long)i << 8) | (unsigned)rc);
            a.classes[i] = -1;
            return;
        }
        a.sdp4[rankDeep] = sdp4_record(ds);
        a.sdp4Orig[rankDeep] = i;
        a.classes[i] = 1 + ds.irez;
    }
}

}  // namespace

uint32_t ingest_block_count(uint32_t n) { return (n + kIngestThreads - 1) / kIngestThreads; }

cudaError_t launch_ingest_classify(const IngestArgs &a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    const uint32_t blocks = ingest_block_count(a.n);
    ingest_classify_kernel<<<blocks, kIngestThreads, 0, stream>>>(a);
    ingest_offsets_kernel<<<1, 1024, 0, stream>>>(a, blocks);
    return cudaGetLastError();
}

cudaError_t launch_ingest_build(const IngestArgs &a, cudaStream_t stream) {
    if (a.n == 0) return cudaSuccess;
    ingest_build_kernel<<<ingest_block_count(a.n), kIngestThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace az
