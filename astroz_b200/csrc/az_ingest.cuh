// az_ingest.cuh -- launch interface of the device-side element initialisation (K5, az_ingest.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "az_tables.hpp"

namespace az {

// All pointers are device pointers.  Two phases with one host read between them, because the size of the tile
// table depends on how many element sets classify as near earth:
//   launch_ingest_classify -> read totals[0..1] and firstFail -> allocate tiles / records -> launch_ingest_build
struct IngestArgs {
    // mean elements, one column each (the fields of src/Tle.zig:8-29 the propagators read)
    const double *epochJd = nullptr, *revPerDay = nullptr, *ecc = nullptr, *inclDeg = nullptr, *raanDeg = nullptr,
                 *argpDeg = nullptr, *maDeg = nullptr, *bstar = nullptr;
    uint32_t n = 0;
    Gravity grav{};
    // phase 1 outputs
    uint8_t *flags = nullptr;             // [n] 0 near earth, 1 deep space, 2 failed
    uint32_t *blockNear = nullptr;        // [ingest_block_count(n)] counts, then exclusive offsets
    uint32_t *blockDeep = nullptr;
    uint32_t *totals = nullptr;           // [2] nSgp4, nSdp4
    unsigned long long *firstFail = nullptr;  // (catalog index << 8 | status), ~0 when every set initialised
    // phase 2 outputs
    double *tiles = nullptr;              // [tiles][kSgp4Cols][8]
    uint32_t *sgp4Orig = nullptr;         // padded
    uint32_t *identity = nullptr;         // padded
    Sdp4Sat *sdp4 = nullptr;              // [nSdp4]
    uint32_t *sdp4Orig = nullptr;
    int32_t *classes = nullptr;           // [n] 0 SGP4, 1 + irez for SDP4
};

uint32_t ingest_block_count(uint32_t n);
cudaError_t launch_ingest_classify(const IngestArgs &a, cudaStream_t stream);
cudaError_t launch_ingest_build(const IngestArgs &a, cudaStream_t stream);

}  // namespace az
