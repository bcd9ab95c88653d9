// az_kernels.cuh -- launch interface of the grid kernels (internal to the library).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "az_device.cuh"

namespace az {

// Arguments of one (n_sats x n_times) grid launch.  All pointers are device pointers.
struct GridArgs {
    // near-earth table (K1) or deep-space records (K2)
    const double *sgp4Tiles = nullptr;   // [tiles][kSgp4Cols][8]
    const Sdp4Sat *sdp4 = nullptr;       // [nSats]
    const double2 *lattice = nullptr;    // K2: [nSats][2][latticeNodes] (xli, xni) at atime = +-720*k
    int latticeNodes = 0;
    const uint32_t *orig = nullptr;      // output row of each table satellite (padded for K1)
    const uint8_t *mask = nullptr;       // nullable: per output row, 0 = leave that satellite's rows untouched
    uint32_t nSats = 0;                  // real satellites in the table
    // time axis
    const double *tbase = nullptr;       // K1: minutes of each epoch relative to the reference epoch
    const double *toff = nullptr;        // K1: per-satellite (reference - epoch) * 1440, padded
    const double *jdFull = nullptr;      // K2: jd + fr per epoch
    const double *tsince = nullptr;      // K2: if set, minutes since epoch are taken from here instead
    const double *jdArr = nullptr;       // K1t: if set, tsince = ((jd[t] + fr[t]) - epochJd) * 1440 on the device
    const double *frArr = nullptr;
    double epochJd = 0.0;
    const double *gsin = nullptr;        // sin/cos(GMST) per epoch when mode != TEME
    const double *gcos = nullptr;
    uint32_t nTimes = 0;
    uint32_t stripe = 0;                 // K1: epochs per CTA, chosen at launch from the CTA count (0 = the shape's default)
    // outputs
    double *pos = nullptr;
    double *vel = nullptr;               // nullable
    uint8_t *status = nullptr;           // nullable, [outRow][nTimes]
    uint32_t outNumSats = 0;             // row count of the output block (time-major stride)
    uint32_t recStride = 3;              // K1t only: doubles between consecutive epochs (6 = x y z vx vy vz records)
    // fused all-gather (satellite-major only): when gather != 0 the result block is written to every GPU
    // of the box from inside the kernel -- gather 1: one multimem.st per 16 bytes to the NVLS multicast
    // mapping of the symmetric buffer (mcPos/mcVel); gather 2: plain stores to each peer mapping.
    int gather = 0;
    int nPeers = 0;
    double *mcPos = nullptr;
    double *mcVel = nullptr;
    double *peerPos[8] = {};
    double *peerVel[8] = {};
    GravConsts g{};
};

constexpr int kMaxPeers = 8;

// K1: near-earth grid.  variant selects a tuning configuration (0 = default).
cudaError_t launch_sgp4_grid(const GridArgs &a, int mode, int layout, cudaStream_t stream, int variant);
// K2a: resonance lattice pre-pass (one thread per deep-space satellite, sequential 720-min steps).
cudaError_t launch_sdp4_lattice(const Sdp4Sat *sats, uint32_t nSats, double2 *lattice, int nodes, cudaStream_t stream);
// K2: deep-space grid.
cudaError_t launch_sdp4_grid(const GridArgs &a, int mode, int layout, cudaStream_t stream);
// K3: fused propagate + single-target conjunction screen (src/Constellation.zig:683-756).  No position block
// is written: per satellite the minimum distance to the target over all epochs (and its epoch index).
struct ScreenArgs {
    const double *sgp4Tiles = nullptr;
    const double *toff = nullptr;     // per-satellite epoch offsets (padded)
    const double *tbase = nullptr;    // times[n_times]
    uint32_t nSats = 0, nTimes = 0;
    uint32_t targetIdx = 0;
    double thresholdSq = 0.0;
    double *track = nullptr;          // scratch [n_times][3]: target positions
    double *minDist = nullptr;        // out [nSats]
    uint32_t *minT = nullptr;         // out [nSats]
    GravConsts g{};
};
cudaError_t launch_sgp4_screen(const ScreenArgs &a, cudaStream_t stream);

// K4: all-vs-all coarse conjunction screen over a device-resident position block
// (bindings/python/src/conjunction.zig:11-149): per epoch a cell list (cell edge = threshold) in a hash table,
// then each satellite checks its own and 13 forward neighbour cells.
struct CoarseArgs {
    const double *pos = nullptr;       // [nSats][nTimes][3] (layout 0) or [nTimes][nSats][3] (layout 1)
    const uint8_t *validMask = nullptr;  // nullable, per satellite
    uint32_t nSats = 0, nTimes = 0;
    int layout = 1;
    double threshold = 0.0;
    uint32_t t0 = 0, tCount = 0;       // epoch batch handled by this launch pair
    uint32_t tableBits = 16;
    uint32_t *head = nullptr;          // [tCount][1 << tableBits]
    uint32_t *next = nullptr;          // [tCount][nSats]
    uint32_t *pairs = nullptr;         // [maxResults][2]
    uint32_t *tIdx = nullptr;          // [maxResults]
    uint32_t maxResults = 0;
    unsigned long long *count = nullptr;  // device counter (total hits, may exceed maxResults)
};
cudaError_t launch_coarse_screen(const CoarseArgs &a, cudaStream_t stream);

// fp32 study kernel (BASELINE config 5): same grid / layout as K1 (satellite-major TEME, fp64 output words),
// arithmetic in fp32; phase64 != 0 forms the secular angles in fp64 first.
cudaError_t launch_sgp4_grid_f32(const GridArgs &a, int phase64, cudaStream_t stream);

// DFMA throughput microbenchmark: returns achieved fp64 FLOP/s (FMA = 2).
cudaError_t measure_fp64_peak(double *flops);
// Arithmetic peak of the fp64 pipe: SMs x 64 lanes x 2 FLOP x the maximum SM clock.
cudaError_t fp64_pipe_peak(double *flops);

int sgp4_variant_count();
void set_sdp4_variant(int v);
void set_sgp4_stripe(uint32_t epochs);  // measurement only: fixed epochs per CTA for the near-earth grid, 0 = automatic
const char *sgp4_variant_name(int variant);

}  // namespace az
