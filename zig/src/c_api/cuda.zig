//! CUDA propagation library bindings -- GENERATED from include/astroz_b200.h by tools/gen_zig_bindings.py.
//! Drop this file in as src/c_api/cuda.zig of ATTron/astroz (next to src/c_api/sgp4.zig); INTEGRATION.md has
//! the build.zig wiring and zig/src/Constellation.device.zig the device branch of Constellation.zig.
//! Error codes are err.Code values (src/c_api/error.zig:3-19) extended with cudaError = -200, noCudaDevice = -201.
//! Uncompiled here: the build image has no Zig toolchain (DESIGN.md section 1).

pub const Handle = ?*anyopaque;

pub extern fn astroz_cuda_version() u32;
pub extern fn astroz_cuda_device_count() i32;
pub extern fn astroz_cuda_last_error() [*:0]const u8;
pub extern fn astroz_cuda_host_alloc(bytes: usize) ?*anyopaque;
pub extern fn astroz_cuda_host_free(p: ?*anyopaque) void;
pub extern fn astroz_cuda_host_register(p: ?*anyopaque, bytes: usize) i32;
pub extern fn astroz_cuda_host_unregister(p: ?*anyopaque) i32;
pub extern fn astroz_cuda_constellation_create(line1: [*]const [*:0]const u8, line2: [*]const [*:0]const u8, n: u32, grav: i32, device: i32, out: *Handle) i32;
pub extern fn astroz_cuda_constellation_create_from_text(text: [*]const u8, len: usize, grav: i32, device: i32, out: *Handle) i32;
pub extern fn astroz_cuda_constellation_create_from_elements(epoch_jd: ?[*]const f64, mean_motion_rev_day: ?[*]const f64, ecc: ?[*]const f64, incl_deg: ?[*]const f64, raan_deg: ?[*]const f64, argp_deg: ?[*]const f64, ma_deg: ?[*]const f64, bstar: ?[*]const f64, n: u32, grav: i32, device: i32, out: *Handle) i32;
pub extern fn astroz_cuda_constellation_create_from_elements_device(d_epoch_jd: ?[*]const f64, d_mean_motion_rev_day: ?[*]const f64, d_ecc: ?[*]const f64, d_incl_deg: ?[*]const f64, d_raan_deg: ?[*]const f64, d_argp_deg: ?[*]const f64, d_ma_deg: ?[*]const f64, d_bstar: ?[*]const f64, n: u32, grav: i32, device: i32, out: *Handle) i32;
pub extern fn astroz_cuda_constellation_free(h: Handle) void;
pub extern fn astroz_cuda_constellation_counts(h: Handle, n: ?[*]u32, n_sgp4: ?[*]u32, n_sdp4: ?[*]u32) i32;
pub extern fn astroz_cuda_constellation_epochs(h: Handle, epochs: ?[*]f64) i32;
pub extern fn astroz_cuda_constellation_classes(h: Handle, classes: ?[*]i32) i32;
pub extern fn astroz_cuda_constellation_get_reference_epoch(h: Handle, jd: ?[*]f64) i32;
pub extern fn astroz_cuda_constellation_set_reference_epoch(h: Handle, jd: f64) i32;
pub extern fn astroz_cuda_constellation_propagate(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, pos: ?[*]f64, vel: ?[*]f64, mode: i32, layout: i32) i32;
pub extern fn astroz_cuda_constellation_propagate_device(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, d_pos: ?[*]f64, d_vel: ?[*]f64, d_status: ?[*]u8, mode: i32, layout: i32, out_num_sats: u32, out_sat_offset: u32, stream: ?*anyopaque) i32;
pub extern fn astroz_cuda_constellation_propagate_gather(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, peer_pos: ?[*]const ?*anyopaque, peer_vel: ?[*]const ?*anyopaque, n_peers: u32, mc_pos: ?*anyopaque, mc_vel: ?*anyopaque, out_num_sats: u32, out_sat_offset: u32, stream: ?*anyopaque) i32;
pub extern fn astroz_cuda_constellation_host_block(h: Handle, n_times: u32, layout: i32, out: ?[*]?[*]f64) i32;
pub extern fn astroz_cuda_constellation_devices(h: Handle, n_devices: ?[*]i32, device_ids: ?[*]i32, first_rows: ?[*]u32) i32;
pub extern fn astroz_cuda_constellation_propagate_replicated(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, velocities: i32, d_pos: ?[*]?[*]f64, d_vel: ?[*]?[*]f64) i32;
pub extern fn astroz_cuda_constellation_reset_carry(h: Handle) i32;
pub extern fn astroz_cuda_sgp4_propagate_into(h: Handle, times: ?[*]const f64, n_times: u32, epoch_offsets: ?[*]const f64, pos: ?[*]f64, vel: ?[*]f64, mode: i32, reference_jd: f64, layout: i32, satellite_mask: ?[*]const u8, out_num_sats: u32) i32;
pub extern fn astroz_cuda_sgp4_propagate_into_device(h: Handle, times: ?[*]const f64, n_times: u32, epoch_offsets: ?[*]const f64, d_pos: ?[*]f64, d_vel: ?[*]f64, mode: i32, reference_jd: f64, layout: i32, satellite_mask: ?[*]const u8, out_num_sats: u32, stream: ?*anyopaque) i32;
pub extern fn astroz_cuda_sdp4_propagate_into(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, pos: ?[*]f64, vel: ?[*]f64, mode: i32, layout: i32, out_num_sats: u32, sat_offset: u32) i32;
pub extern fn astroz_cuda_sdp4_propagate_into_device(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, d_pos: ?[*]f64, d_vel: ?[*]f64, mode: i32, layout: i32, out_num_sats: u32, sat_offset: u32, stream: ?*anyopaque) i32;
pub extern fn astroz_cuda_sgp4_screen(h: Handle, times: ?[*]const f64, n_times: u32, epoch_offsets: ?[*]const f64, target_idx: u32, threshold: f64, reference_jd: f64, out_min_dists: ?[*]f64, out_min_t: ?[*]u32) i32;
pub extern fn astroz_cuda_constellation_coarse_screen_device(h: Handle, d_positions: ?[*]const f64, num_sats: u32, num_times: u32, layout: i32, threshold: f64, d_valid_mask: ?[*]const u8, d_pairs: ?[*]u32, d_t_indices: ?[*]u32, max_results: u32, count: *u64) i32;
pub extern fn astroz_cuda_sgp4_screen_all(h: Handle, times: ?[*]const f64, n_times: u32, epoch_offsets: ?[*]const f64, threshold: f64, pairs: ?[*]u32, t_indices: ?[*]u32, max_results: u32, count: *u64) i32;
pub extern fn astroz_cuda_constellation_synchronize(h: Handle) i32;
pub extern fn astroz_cuda_constellation_set_timing(h: Handle, enabled: i32) i32;
pub extern fn astroz_cuda_constellation_last_kernel_ms(h: Handle, ms: *[3]f32) i32;
pub extern fn astroz_cuda_sgp4_init(line1: [*:0]const u8, line2: [*:0]const u8, grav: i32, device: i32, out: *Handle) i32;
pub extern fn astroz_cuda_sgp4_free(h: Handle) void;
pub extern fn astroz_cuda_sgp4_is_deep_space(h: Handle) i32;
pub extern fn astroz_cuda_sgp4_epoch(h: Handle, epoch_jd: ?[*]f64) i32;
pub extern fn astroz_cuda_sgp4_elements(h: Handle, out10: ?[*]f64) i32;
pub extern fn astroz_cuda_sgp4_propagate(h: Handle, tsince: f64, pos: *[3]f64, vel: *[3]f64) i32;
pub extern fn astroz_cuda_sgp4_propagate_batch(h: Handle, times: ?[*]const f64, results: ?[*]f64, count: u32) i32;
pub extern fn astroz_cuda_sgp4_array(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, epoch_jd: f64, results: ?[*]f64, count: u32) i32;
pub extern fn astroz_cuda_constellation_propagate_device_f32(h: Handle, jd: ?[*]const f64, fr: ?[*]const f64, n_times: u32, d_pos: ?[*]f64, d_vel: ?[*]f64, phase64: i32, stream: ?*anyopaque) i32;
pub extern fn astroz_cuda_fp64_peak(device: i32, tflops: ?[*]f64) i32;
pub extern fn astroz_cuda_fp64_pipe_peak(device: i32, tflops: ?[*]f64) i32;

/// C API code -> the error set of the kernel-level boundary it replaces (src/simdKernels.zig:30-37)
pub fn toError(rc: i32) ?@import("../Sgp4.zig").Error {
    return switch (rc) {
        0 => null,
        -12 => error.SatelliteDecayed,
        -11 => error.InvalidEccentricity,
        -10 => error.DeepSpaceNotSupported,
        else => error.OutOfMemory, // -100 alloc, -200 CUDA, -201 no device: no CPU fallback is attempted
    };
}
