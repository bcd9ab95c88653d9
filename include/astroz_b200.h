/*
 * astroz_b200.h -- C ABI of the B200-native batch SGP4/SDP4 propagator.
 *
 * Drop-in boundary for the propagation path of ATTron/astroz (reference paths relative to the
 * reference repo root).  Every entry point names the reference interface it replaces.
 * Plain pointers and sizes only; no torch / C++ types.  All functions return an int32 from the
 * reference's own error-code space (src/c_api/error.zig:3-19) extended with CUDA codes.
 *
 * Threading / ownership (same contract as the reference, src/c_api/allocator.zig:9-18 and
 * src/Constellation.zig:88,294): handles are not thread-safe; caller owns every I/O buffer and the
 * library never retains it; device memory is owned by the handle and released by the matching *_free.
 * A handle keeps ONE copy of its per-call device scratch (time axis, epoch offsets, mask): calls on a handle
 * must be issued on one stream, or the caller synchronises between calls that use different streams.
 *
 * Failed cells.  Deep-space cells are checked like the reference's scalar path (src/Sdp4.zig:913-967: mean
 * motion <= 0, eccentricity >= 1 or < -0.001, semi-major axis < 0.95, radius < 1 earth radius) and a failing
 * cell is zero-filled, per satellite (the reference's batch path zero-fills the 8 satellites of the batch,
 * src/Constellation.zig:468-471,511-528, and applies only the first three checks, src/Sdp4Batch.zig:293-324).
 * Near-earth cells are never zero-filled -- the reference's near-earth batch path has no reachable failure
 * (src/Sgp4Batch.zig:147-150) -- the state is stored and the optional status byte reports radius < 1 earth
 * radius as ASTROZ_CELL_DECAYED (the scalar path's check, src/Sgp4.zig:588-590) as a diagnostic.
 */
#ifndef ASTROZ_B200_H
#define ASTROZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes: src/c_api/error.zig:3-19 (+ CUDA extensions <= -200) ------------------------- */
#define ASTROZ_OK                    0
#define ASTROZ_BAD_TLE_LENGTH      (-1)
#define ASTROZ_BAD_CHECKSUM        (-2)
#define ASTROZ_DEEP_SPACE          (-10)   /* deepSpaceNotSupported */
#define ASTROZ_INVALID_ECC         (-11)
#define ASTROZ_DECAYED             (-12)
#define ASTROZ_VALUE_ERROR         (-20)
#define ASTROZ_ALLOC_FAILED        (-100)
#define ASTROZ_NULL_POINTER        (-101)
#define ASTROZ_NOT_INITIALIZED     (-102)
#define ASTROZ_UNKNOWN             (-999)
#define ASTROZ_CUDA_ERROR          (-200)  /* any CUDA runtime failure; see astroz_cuda_last_error() */
#define ASTROZ_NO_DEVICE           (-201)  /* no CUDA device: the library never falls back to a CPU path */

/* ---- per-cell status bytes (kernel-level codes, src/simdKernels.zig:30-37) --------------------- */
#define ASTROZ_CELL_OK           0
#define ASTROZ_CELL_DECAYED      1
#define ASTROZ_CELL_INVALID_ECC  2

/* gravity model selector: src/c_api/sgp4.zig:17-20 (0 = WGS84, 1 = WGS72) */
#define ASTROZ_WGS84 0
#define ASTROZ_WGS72 1

/* src/Constellation.zig:30-42 */
#define ASTROZ_MODE_TEME      0
#define ASTROZ_MODE_ECEF      1
#define ASTROZ_MODE_GEODETIC  2
#define ASTROZ_LAYOUT_SATELLITE_MAJOR 0   /* (n_sats, n_times, 3) */
#define ASTROZ_LAYOUT_TIME_MAJOR      1   /* (n_times, n_sats, 3) */

typedef void *astroz_constellation_t;
typedef void *astroz_sgp4_t;

/* replaces astroz_version (src/c_api/root.zig:13-15): (major<<16)|(minor<<8)|patch */
uint32_t astroz_cuda_version(void);
/* number of visible CUDA devices (0 when none) */
int32_t astroz_cuda_device_count(void);
/* text of the last CUDA failure seen on this thread ("" if none); pointer valid until the next call */
const char *astroz_cuda_last_error(void);

/* pinned host buffers for zero-staging device<->host copies of the output block */
void *astroz_cuda_host_alloc(size_t bytes);
void astroz_cuda_host_free(void *p);
/* Caller-owned buffers.  The host-buffer calls accept ANY host memory.  A buffer from astroz_cuda_host_alloc, or
 * one page-locked with astroz_cuda_host_register (cudaHostRegister: costs about as much as touching the pages once,
 * so it pays for blocks that are reused), receives the result by direct DMA.  Plain pageable memory (a numpy array, a
 * Zig slice from the page allocator) is served through a ring of pinned slots inside the handle: the result leaves the
 * GPU in 32 MB pieces at the full PCIe rate and a small pool of host threads (ASTROZ_COPY_THREADS, default 12, streaming stores) copies
 * each landed piece to its place while the next ones are in flight. */
int32_t astroz_cuda_host_register(void *p, size_t bytes);
int32_t astroz_cuda_host_unregister(void *p);

/* ------------------------------------------------------------------------------------------------
 * Constellation: replaces Constellation.init / propagate / resetCarry / deinit
 * (src/Constellation.zig:101-200, 245-308, 214-218, 202-210).
 * ---------------------------------------------------------------------------------------------- */

/* Parse n TLEs (NUL-terminated 69-column lines, src/Tle.zig:49-101), classify each as SGP4 or SDP4
 * exactly as src/Constellation.zig:115-126, build the device element tables on `device`.
 * device = -1: a MULTI-DEVICE handle -- the catalog is cut into contiguous, 8-row-aligned satellite ranges of
 * equal cost, one per visible GPU (the analogue of the reference's thread fan-out over one propagate call,
 * src/Constellation.zig:327-385).  The environment variable ASTROZ_DEVICES caps the number of GPUs used, the way
 * ASTROZ_THREADS caps the reference's threads (src/Constellation.zig:61-74); ASTROZ_DEVICE_LIST="0,2,3" names
 * ordinals explicitly.  The host-buffer calls (astroz_cuda_constellation_propagate, astroz_cuda_sgp4_propagate_into)
 * then run every GPU at once, each copying its rows over its own PCIe link into its slice of the caller's block;
 * results are bit-identical to a single-device handle.  Entry points that take DEVICE pointers need a
 * single-device handle and return ASTROZ_VALUE_ERROR otherwise.  (The three host constructors accept -1.)
 * Errors: BAD_TLE_LENGTH, INVALID_ECC, DECAYED (first offending TLE aborts, as the reference). */
int32_t astroz_cuda_constellation_create(const char *const *line1, const char *const *line2, uint32_t n,
                                         int32_t grav, int32_t device, astroz_constellation_t *out);

/* Same, from a text blob of 2- or 3-line element sets (src/Tle.zig:103-132 MultiIterator semantics). */
int32_t astroz_cuda_constellation_create_from_text(const char *text, size_t len, int32_t grav, int32_t device,
                                                   astroz_constellation_t *out);

/* Same, from numeric mean elements (the fields Tle.parseOmm fills from an OMM record, src/Tle.zig:134-215):
 * epoch_jd, mean motion [rev/day], eccentricity, inclination / RAAN / argument of perigee / mean anomaly [deg],
 * B* [1/earth radii].  No text round trip, so Monte-Carlo draws keep their full fp64 values. */
int32_t astroz_cuda_constellation_create_from_elements(const double *epoch_jd, const double *mean_motion_rev_day,
                                                       const double *ecc, const double *incl_deg, const double *raan_deg,
                                                       const double *argp_deg, const double *ma_deg, const double *bstar,
                                                       uint32_t n, int32_t grav, int32_t device,
                                                       astroz_constellation_t *out);

/* Same, with the eight element columns already resident in HBM on `device` (DEVICE pointers): classification,
 * Sgp4.initElements / Sdp4.initElements (src/Sgp4.zig:108-417, src/Sdp4.zig:174-657) and the table build of
 * src/Constellation.zig:101-200 run on the GPU, one element set per thread; nothing but counts, epochs and the row
 * maps (16 B per set) returns to the host.  For Monte-Carlo draws and OMM streams generated on the device. */
int32_t astroz_cuda_constellation_create_from_elements_device(
    const double *d_epoch_jd, const double *d_mean_motion_rev_day, const double *d_ecc, const double *d_incl_deg,
    const double *d_raan_deg, const double *d_argp_deg, const double *d_ma_deg, const double *d_bstar, uint32_t n,
    int32_t grav, int32_t device, astroz_constellation_t *out);

void astroz_cuda_constellation_free(astroz_constellation_t h);

/* numSatellites / numSgp4 / numSdp4 (src/Constellation.zig:82,89,95) */
int32_t astroz_cuda_constellation_counts(astroz_constellation_t h, uint32_t *n, uint32_t *n_sgp4, uint32_t *n_sdp4);
/* per-satellite epoch JD (n doubles) and class (n int32: 0 SGP4, 1 SDP4 irez0, 2 irez1, 3 irez2) */
int32_t astroz_cuda_constellation_epochs(astroz_constellation_t h, double *epochs);
int32_t astroz_cuda_constellation_classes(astroz_constellation_t h, int32_t *classes);
/* referenceEpochJd (src/Constellation.zig:92,139-140); settable so satellite shards of one catalog
 * share the whole catalog's reference epoch */
int32_t astroz_cuda_constellation_get_reference_epoch(astroz_constellation_t h, double *jd);
int32_t astroz_cuda_constellation_set_reference_epoch(astroz_constellation_t h, double jd);

/* Constellation.propagate (src/Constellation.zig:245-308) with HOST buffers.
 * pos / vel: n*n_times*3 doubles each (vel may be NULL), written per `layout`; `mode` selects TEME/ECEF/geodetic.
 * Cells whose propagation fails are zero-filled (src/Constellation.zig:511-528).  Includes the
 * host->device copy of jd/fr and the device->host copy of the result block. */
int32_t astroz_cuda_constellation_propagate(astroz_constellation_t h, const double *jd, const double *fr,
                                            uint32_t n_times, double *pos, double *vel, int32_t mode, int32_t layout);

/* Same computation, results left in HBM (d_pos / d_vel are DEVICE pointers on the handle's device).
 * out_num_sats / out_sat_offset place this handle's satellites inside a larger output block
 * (the reference's numSatellites-as-stride convention, src/Constellation.zig:46-51 and
 * bindings/python/src/sgp4.zig:216); pass n and 0 for a stand-alone constellation.
 * d_status (nullable): n*n_times bytes, satellite-major, per-cell ASTROZ_CELL_* code.
 * stream: a cudaStream_t (NULL = the handle's own stream); the call is asynchronous on it. */
int32_t astroz_cuda_constellation_propagate_device(astroz_constellation_t h, const double *jd, const double *fr,
                                                   uint32_t n_times, double *d_pos, double *d_vel, uint8_t *d_status,
                                                   int32_t mode, int32_t layout, uint32_t out_num_sats,
                                                   uint32_t out_sat_offset, void *stream);

/* Fused propagate + all-gather for satellite-sharded multi-GPU runs (no reference
 * counterpart -- the reference is single-process).  TEME, satellite-major.  This handle's rows
 * [out_sat_offset, out_sat_offset + n) of the (out_num_sats, n_times, 3) block are written, from inside the
 * propagation kernels, into EVERY GPU's copy of the block:
 *   mc_pos / mc_vel != NULL : NVLS multicast mappings of a symmetric allocation -- one multimem.st per
 *                             16 bytes, NVSwitch replicates it to all GPUs (this one included);
 *   otherwise               : peer_pos[0..n_peers) / peer_vel[...] are the per-GPU mappings of the block
 *                             (this GPU's own mapping included) and each run is stored to all of them.
 * peer_vel / mc_vel may be NULL (positions only).  Asynchronous on `stream`; the caller synchronises the
 * ranks (e.g. a symmetric-memory barrier) before reading other ranks' rows. */
int32_t astroz_cuda_constellation_propagate_gather(astroz_constellation_t h, const double *jd, const double *fr,
                                                   uint32_t n_times, void *const *peer_pos, void *const *peer_vel,
                                                   uint32_t n_peers, void *mc_pos, void *mc_vel,
                                                   uint32_t out_num_sats, uint32_t out_sat_offset, void *stream);

/* A page-locked result block of n * n_times * 3 doubles placed for the handle that will fill it: on a multi-device
 * handle each device's satellite range of a satellite-major block is bound to the NUMA node that device hangs off
 * (mbind), so every GPU writes node-local host memory over its own PCIe link -- a plain pinned allocation lives on one
 * node and was measured at 93 GB/s for 8 GPUs against 315 GB/s node-local.  Free with astroz_cuda_host_free. */
int32_t astroz_cuda_constellation_host_block(astroz_constellation_t h, uint32_t n_times, int32_t layout, double **out);

/* Devices behind a handle: *n_devices (1 for a single-device handle); device_ids[n_devices] (nullable) their CUDA
 * ordinals; first_rows[n_devices + 1] (nullable) the first catalog row of each device's satellite range, then n. */
int32_t astroz_cuda_constellation_devices(astroz_constellation_t h, int32_t *n_devices, int32_t *device_ids,
                                          uint32_t *first_rows);

/* The north star's all-gather behind one handle, without an external communicator: propagate (TEME,
 * satellite-major) and leave the WHOLE (n, n_times, 3) position block -- and velocity block when velocities != 0 --
 * in the HBM of EVERY device of a multi-device handle.  Each device's propagation kernels store their rows, run by
 * run, straight into all devices' copies over NVLink (peer mappings from cudaDeviceEnablePeerAccess), so the
 * transfer overlaps the fp64 work; the call returns when every copy is complete.  d_pos[k] / d_vel[k] receive the
 * block pointers on device k (owned by the handle, valid until the next call or free); arrays of n_devices entries.
 * On a single-device handle the block is simply left on that device. */
int32_t astroz_cuda_constellation_propagate_replicated(astroz_constellation_t h, const double *jd, const double *fr,
                                                       uint32_t n_times, int32_t velocities, double **d_pos,
                                                       double **d_vel);

/* Constellation.resetCarry (src/Constellation.zig:214-218).  The device path re-derives the SDP4
 * resonance state from a 720-minute lattice on every call, so this is a semantic no-op kept for drop-in use. */
int32_t astroz_cuda_constellation_reset_carry(astroz_constellation_t h);

/* stateless near-earth path: replaces Constellation.propagateConstellation (src/Constellation.zig:541-605)
 * as called by SatrecArray.propagate_into / Sgp4Constellation.propagate_into
 * (bindings/python/src/satrec.zig:896-988, bindings/python/src/sgp4.zig:171-268):
 *   tsince[sat][t] = times[t] + epoch_offsets[sat]   (minutes)
 * Only the SGP4 satellites of `h` take part, in catalog order: satellite i -> output row i.
 * epoch_offsets: n_sgp4 doubles.  reference_jd is used for GMST when mode != TEME.
 * satellite_mask (nullable, n_sgp4 bytes): rows whose byte is 0 are not computed and not written
 * (src/Constellation.zig:436-446,530-533).  out_num_sats (0 = n_sgp4): row count of the output block, the
 * reference's output_stride (bindings/python/src/sgp4.zig:215-216).  HOST buffers of out_num_sats*n_times*3. */
int32_t astroz_cuda_sgp4_propagate_into(astroz_constellation_t h, const double *times, uint32_t n_times,
                                        const double *epoch_offsets, double *pos, double *vel, int32_t mode,
                                        double reference_jd, int32_t layout, const uint8_t *satellite_mask,
                                        uint32_t out_num_sats);
/* device-resident variant of the above (d_pos/d_vel device pointers; the mask is still a HOST array) */
int32_t astroz_cuda_sgp4_propagate_into_device(astroz_constellation_t h, const double *times, uint32_t n_times,
                                               const double *epoch_offsets, double *d_pos, double *d_vel,
                                               int32_t mode, double reference_jd, int32_t layout,
                                               const uint8_t *satellite_mask, uint32_t out_num_sats, void *stream);

/* stateless deep-space path: replaces Constellation.propagateSdp4Constellation (src/Constellation.zig:611-674) as
 * called by sdp4_batch_propagate_into (bindings/python/src/satrec.zig:505-644).  Only the SDP4 satellites of `h` take
 * part, in catalog order: deep-space satellite i -> output row sat_offset + i of a block with out_num_sats rows
 * (0 = numSdp4), tsince = (jd[t] + fr[t] - epoch) * 1440 (src/Sdp4Batch.zig:199-215).  HOST buffers; rows of other
 * satellites are not touched. */
int32_t astroz_cuda_sdp4_propagate_into(astroz_constellation_t h, const double *jd, const double *fr, uint32_t n_times,
                                        double *pos, double *vel, int32_t mode, int32_t layout, uint32_t out_num_sats,
                                        uint32_t sat_offset);
/* device-resident variant (d_pos / d_vel DEVICE pointers to the whole out_num_sats-row block) */
int32_t astroz_cuda_sdp4_propagate_into_device(astroz_constellation_t h, const double *jd, const double *fr,
                                               uint32_t n_times, double *d_pos, double *d_vel, int32_t mode,
                                               int32_t layout, uint32_t out_num_sats, uint32_t sat_offset, void *stream);

/* Fused propagate + single-target conjunction screen: replaces Constellation.screenConstellation
 * (src/Constellation.zig:683-756) as called by Sgp4Constellation.screen_conjunction
 * (bindings/python/src/sgp4.zig).  Near-earth satellites only, tsince = times[t] + epoch_offsets[sat].
 * out_min_dists[n_sgp4]: minimum distance (km) to satellite `target_idx` over all times, or `threshold`
 * when never closer; out_min_t[n_sgp4]: index of the (first) time of that minimum, 0 when none.  The
 * target's own entry stays (threshold, 0).  No position block is produced or copied: 12 bytes per
 * satellite come back.  reference_jd is accepted for signature parity (a common GMST rotation does not
 * change distances).  HOST buffers. */
int32_t astroz_cuda_sgp4_screen(astroz_constellation_t h, const double *times, uint32_t n_times,
                                const double *epoch_offsets, uint32_t target_idx, double threshold,
                                double reference_jd, double *out_min_dists, uint32_t *out_min_t);

/* All-vs-all coarse conjunction screen: replaces coarse_screen / coarseScreen
 * (bindings/python/src/conjunction.zig:11-149).  d_positions: DEVICE block (num_sats, num_times, 3) for
 * layout 0 or (num_times, num_sats, 3) for layout 1 (e.g. left in HBM by *_propagate_device);
 * d_valid_mask: nullable per-satellite bytes.  Every (s, other, t) with s < other closer than `threshold`
 * at epoch t is appended to d_pairs[2k], d_pairs[2k+1], d_t_indices[k] (device buffers of max_results
 * entries; order unspecified -- the reference emits the same set ordered by epoch).  *count receives the
 * number of hits found, which may exceed max_results (then only max_results were stored).  Synchronous. */
int32_t astroz_cuda_constellation_coarse_screen_device(astroz_constellation_t h, const double *d_positions,
                                                       uint32_t num_sats, uint32_t num_times, int32_t layout,
                                                       double threshold, const uint8_t *d_valid_mask, uint32_t *d_pairs,
                                                       uint32_t *d_t_indices, uint32_t max_results, uint64_t *count);

/* The all-vs-all branch of astroz.screen(source, times, threshold) (bindings/python/astroz/__init__.py:535-650):
 * propagate the near-earth satellites (tsince = times[t] + epoch_offsets[sat], TEME, positions only), keep the
 * block in HBM, run the coarse screen on it, return only the hits.  pairs / t_indices: HOST buffers of
 * max_results entries. */
int32_t astroz_cuda_sgp4_screen_all(astroz_constellation_t h, const double *times, uint32_t n_times,
                                    const double *epoch_offsets, double threshold, uint32_t *pairs, uint32_t *t_indices,
                                    uint32_t max_results, uint64_t *count);

/* block until everything queued on the handle's stream has finished */
int32_t astroz_cuda_constellation_synchronize(astroz_constellation_t h);

/* Kernel timing is opt-in: enabled != 0 makes every later propagate call on the handle record CUDA events around its
 * kernels (about a dozen microseconds of stream time per call, which is why it is off by default; the environment
 * variable ASTROZ_TIMING=1 turns it on for new handles).  astroz_cuda_constellation_last_kernel_ms returns
 * ASTROZ_NOT_INITIALIZED for a call made with timing off. */
int32_t astroz_cuda_constellation_set_timing(astroz_constellation_t h, int32_t enabled);

/* device time (ms, CUDA events on the launching stream) of the propagation kernels of the last
 * propagate call on this handle: [0] SGP4 grid kernel, [1] span of all grid launches of the call (the two grids of a
 * mixed catalog run side by side on two streams, so [1] < [0] + [2] there), [2] SDP4 grid kernel.
 * After a HOST-buffer call (whose grid is launched in chunks so copies overlap compute) all three slots hold the span
 * from the first kernel of the first chunk to the last kernel of the last chunk; on a multi-device handle, the
 * maximum over its devices. */
int32_t astroz_cuda_constellation_last_kernel_ms(astroz_constellation_t h, float ms[3]);

/* ------------------------------------------------------------------------------------------------
 * Single satellite: replaces sgp4_init / sgp4_free / sgp4_propagate / sgp4_propagate_batch
 * (src/c_api/root.zig:48-59, src/c_api/sgp4.zig:16-100) -- extended to deep-space objects the way
 * Satrec.twoline2rv falls back to SDP4 (bindings/python/src/satrec.zig:135-160).
 * ---------------------------------------------------------------------------------------------- */
int32_t astroz_cuda_sgp4_init(const char *line1, const char *line2, int32_t grav, int32_t device, astroz_sgp4_t *out);
void astroz_cuda_sgp4_free(astroz_sgp4_t h);
/* 1 if the object is propagated with SDP4 (Satrec.is_deep_space) */
int32_t astroz_cuda_sgp4_is_deep_space(astroz_sgp4_t h);
int32_t astroz_cuda_sgp4_epoch(astroz_sgp4_t h, double *epoch_jd);
/* Mean elements behind the python-sgp4 attribute getters of Satrec (bindings/python/src/satrec.zig:395-470):
 * out[10] = ecco, inclo, nodeo, argpo, mo (rad), no_kozai (rad/min), bstar, a (un-Kozai'd semi-major axis, earth
 * radii: Sgp4.Elements.a, src/Sgp4.zig:206-228), no_unkozai (rad/min), epoch JD. */
int32_t astroz_cuda_sgp4_elements(astroz_sgp4_t h, double *out10);
/* one time: pos[3] km, vel[3] km/s (TEME) */
int32_t astroz_cuda_sgp4_propagate(astroz_sgp4_t h, double tsince, double pos[3], double vel[3]);
/* count times (minutes since epoch); results[count][6] = x y z vx vy vz (src/c_api/sgp4.zig:60-100) */
int32_t astroz_cuda_sgp4_propagate_batch(astroz_sgp4_t h, const double *times, double *results, uint32_t count);

/* Satrec.sgp4_array_into(jd, fr, positions, velocities) (bindings/python/src/satrec.zig:299-343): count absolute
 * epochs jd[i] + fr[i]; tsince = ((jd + fr) - epoch_jd) * 1440 is formed on the device in fp64 (the same
 * expression, :263).  results[count][6] = x y z vx vy vz.  Near-earth objects run on the time-parallel kernel. */
int32_t astroz_cuda_sgp4_array(astroz_sgp4_t h, const double *jd, const double *fr, double epoch_jd, double *results,
                               uint32_t count);

/* BASELINE config 5 (fp32 tolerance study; defined by this project, the reference has no such path): the
 * near-earth satellites of `h` propagated with the arithmetic in fp32 -- phase64 = 0: everything in fp32;
 * phase64 = 1: the three secular angles formed and reduced in fp64, the rest in fp32.  Satellite-major TEME,
 * results widened to fp64 words so they can be compared with astroz_cuda_constellation_propagate_device. */
int32_t astroz_cuda_constellation_propagate_device_f32(astroz_constellation_t h, const double *jd, const double *fr,
                                                       uint32_t n_times, double *d_pos, double *d_vel, int32_t phase64,
                                                       void *stream);

/* ---- measurement helpers --------------------------------------------------------------------- */
/* DFMA microbenchmark on `device`: achieved fp64 TFLOP/s (FMA = 2) -- the measured roofline denominator */
int32_t astroz_cuda_fp64_peak(int32_t device, double *tflops);
/* Arithmetic peak of the fp64 pipe of `device`: SMs x 64 FMA lanes x 2 FLOP x the maximum SM clock, in TFLOP/s.
 * bench.py reports the roofline against the larger of this and the live microbenchmark. */
int32_t astroz_cuda_fp64_pipe_peak(int32_t device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* ASTROZ_B200_H */
