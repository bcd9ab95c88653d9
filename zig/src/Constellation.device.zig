//! Device branch of src/Constellation.zig (ATTron/astroz) -- the changes a maintainer merges into that file to route
//! `Constellation.propagate` and the stateless `propagateConstellation` / `propagateSdp4Constellation` /
//! `screenConstellation` entry points through libastroz_b200.so (include/astroz_b200.h, bound in c_api/cuda.zig).
//! Line numbers refer to the reference tree.  The public surface (`init`, `propagate`, `resetCarry`, `deinit`,
//! src/Constellation.zig:101,245,214,202) is unchanged; the struct gains one optional handle.
//! Uncompiled here: the build image has no Zig toolchain (DESIGN.md section 1); every call below is exercised through
//! the same C ABI by tests/test_gpu_parity.py.

const std = @import("std");
const build_options = @import("build_options");
const cuda = if (build_options.enable_cuda) @import("c_api/cuda.zig") else struct {};
const Sgp4 = @import("Sgp4.zig");
const Error = Sgp4.Error;

// --- new field of `Constellation` (next to numSatellites, :95) -------------------------------------------------------
// cudaHandle: if (build_options.enable_cuda) cuda.Handle else void = if (build_options.enable_cuda) null else {},

/// ASTROZ_DEVICES caps the GPUs a `device = -1` handle spreads over, the way ASTROZ_THREADS caps the worker threads of
/// the CPU path (getMaxThreads, :61-74); the library reads it itself, this helper only mirrors the reference's knob
/// for callers that want to know the fan-out.
pub fn getMaxDevices() usize {
    const visible: usize = @intCast(@max(cuda.astroz_cuda_device_count(), 0));
    if (std.posix.getenv("ASTROZ_DEVICES")) |v| {
        const k = std.fmt.parseInt(usize, v, 10) catch return visible;
        if (k >= 1) return @min(k, visible);
    }
    return visible;
}

/// Attach the device path: `lines1/lines2` are the same TLEs `init` was built from, in the same order.
/// device >= 0: that GPU.  device = -1: every visible GPU (capped by ASTROZ_DEVICES) -- one `propagate` call then fans
/// out over the devices exactly as propagateImpl fans out over threads (:327-385), each GPU copying its satellite range
/// over its own PCIe link into the caller's slices.
pub fn attachCuda(self: anytype, lines1: []const [*:0]const u8, lines2: []const [*:0]const u8, grav: i32, device: i32) Error!void {
    if (!build_options.enable_cuda) return;
    var h: cuda.Handle = null;
    if (cuda.toError(cuda.astroz_cuda_constellation_create(lines1.ptr, lines2.ptr, @intCast(lines1.len), grav, device, &h))) |e| return e;
    // classification must agree with the CPU init above (:115-126)
    var n: u32 = 0;
    var ns: u32 = 0;
    var nd: u32 = 0;
    _ = cuda.astroz_cuda_constellation_counts(h, &n, &ns, &nd);
    std.debug.assert(ns == self.numSgp4 and nd == self.numSdp4);
    _ = cuda.astroz_cuda_constellation_set_reference_epoch(h, self.referenceEpochJd); // :139-140
    self.cudaHandle = h;
}

/// Body to place at the top of `propagate` (:245-258), after the two length checks that stay as they are.
/// Returns true when the device path handled the call.  `resultsPos` / `resultsVel` are the caller's own slices:
/// pageable memory is served through the handle's pinned ring, page-locked memory (astroz_cuda_host_alloc /
/// astroz_cuda_host_register) by direct DMA.
pub fn propagateDevice(self: anytype, jd: []const f64, fr: []const f64, resultsPos: []f64, resultsVel: ?[]f64, outputMode: anytype, layout: anytype) Error!bool {
    if (!build_options.enable_cuda) return false;
    const h = self.cudaHandle orelse return false;
    const rc = cuda.astroz_cuda_constellation_propagate(
        h,
        jd.ptr,
        fr.ptr,
        @intCast(jd.len),
        resultsPos.ptr,
        if (resultsVel) |rv| rv.ptr else null,
        @intFromEnum(outputMode), // teme / ecef / geodetic = 0 / 1 / 2 (:30-34)
        @intFromEnum(layout), // satelliteMajor / timeMajor = 0 / 1 (:37-42)
    );
    if (cuda.toError(rc)) |e| return e; // no CPU fallback on error
    return true;
}

/// `propagateConstellation` (:541-605) with a handle built from the same satellites: tsince = times[t] + epochOffsets[sat],
/// satellite i -> row i of a block with `numSatellites` rows (the reference's stride convention, :46-51), rows whose
/// mask byte is 0 untouched (:436-446,530-533).
pub fn propagateConstellationDevice(h: cuda.Handle, numSatellites: usize, times: []const f64, epochOffsets: []const f64, resultsPos: []f64, resultsVel: ?[]f64, outputMode: anytype, referenceJd: f64, satelliteMask: ?[]const u8, layout: anytype) Error!void {
    const rc = cuda.astroz_cuda_sgp4_propagate_into(
        h,
        times.ptr,
        @intCast(times.len),
        epochOffsets.ptr,
        resultsPos.ptr,
        if (resultsVel) |rv| rv.ptr else null,
        @intFromEnum(outputMode),
        referenceJd,
        @intFromEnum(layout),
        if (satelliteMask) |m| m.ptr else null,
        @intCast(numSatellites),
    );
    if (cuda.toError(rc)) |e| return e;
}

/// `propagateSdp4Constellation` (:611-674; origIndices are always sat_offset + i, bindings/python/src/satrec.zig:628-631).
pub fn propagateSdp4ConstellationDevice(h: cuda.Handle, numSatellites: usize, satOffset: usize, jd: []const f64, fr: []const f64, resultsPos: []f64, resultsVel: ?[]f64, outputMode: anytype, layout: anytype) Error!void {
    const rc = cuda.astroz_cuda_sdp4_propagate_into(
        h,
        jd.ptr,
        fr.ptr,
        @intCast(jd.len),
        resultsPos.ptr,
        if (resultsVel) |rv| rv.ptr else null,
        @intFromEnum(outputMode),
        @intFromEnum(layout),
        @intCast(numSatellites),
        @intCast(satOffset),
    );
    if (cuda.toError(rc)) |e| return e;
}

/// `screenConstellation` (:683-756): minimum distance to one target and its first epoch index per satellite, reduced on
/// the device (targetEpochOffset is epochOffsets[targetIdx], which is what every reference caller passes).
pub fn screenConstellationDevice(h: cuda.Handle, times: []const f64, epochOffsets: []const f64, targetIdx: usize, threshold: f64, referenceJd: f64, outMinDists: []f64, outMinTIndices: []u32) Error!void {
    const rc = cuda.astroz_cuda_sgp4_screen(h, times.ptr, @intCast(times.len), epochOffsets.ptr, @intCast(targetIdx), threshold, referenceJd, outMinDists.ptr, outMinTIndices.ptr);
    if (cuda.toError(rc)) |e| return e;
}

/// `resetCarry` (:214-218): the device path re-derives the resonance state from its 720-minute lattice on every call.
pub fn resetCarryDevice(self: anytype) void {
    if (build_options.enable_cuda) if (self.cudaHandle) |h| {
        _ = cuda.astroz_cuda_constellation_reset_carry(h);
    };
}

/// `deinit` (:202-210): release the handle before the existing frees.
pub fn deinitDevice(self: anytype) void {
    if (build_options.enable_cuda) if (self.cudaHandle) |h| cuda.astroz_cuda_constellation_free(h);
}

// --- the three call sites in Constellation.zig -------------------------------------------------------------------------
// pub fn propagate(self: *Constellation, jd, fr, resultsPos, resultsVel, outputMode, layout) Error!void {
//     ... length checks (:255-257) ...
//     if (try device.propagateDevice(self, jd, fr, resultsPos, resultsVel, outputMode, layout)) return;
//     ... existing CPU path (:259-308) ...
// }
// pub fn resetCarry(self: *Constellation) void { device.resetCarryDevice(self); ... existing loop ... }
// pub fn deinit(self: *Constellation) void { device.deinitDevice(self); ... existing frees ... }
