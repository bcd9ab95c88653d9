// az_device_f32.cuh -- fp32 near-earth cell for the BASELINE config-5 precision study.
//
// There is no reference path for this: astroz propagates in f64 only, and its MonteCarlo.zig is a
// Hohmann delta-v simulator (src/MonteCarlo.zig:93-157).  The study asks what single precision would
// cost.  Two variants: kPhase64 = false computes everything, including the secular angles mo + mdot*t,
// in fp32; kPhase64 = true forms the three secular angles in fp64, reduces them mod 2pi, and runs the
// rest (drag terms, Kepler solve, short-period terms, orientation) in fp32.  The algebra is the plain
// SGP4 formulation of src/Sgp4Batch.zig:113-157 + src/Sgp4.zig:646-750 with CUDA's float intrinsics.
#pragma once

#include "az_device.cuh"

namespace az {

template <bool kPhase64, typename ColFn>
__device__ __forceinline__ void sgp4_cell_f32(ColFn col, double td, const GravConsts &g, CellOut &o) {
    const float t = (float)td;
    const float t2 = t * t;
    float xmdf, argpdf, nodem;
    if (kPhase64) {
        xmdf = (float)mod_twopi(fma(col(kMdot), td, col(kMo)));
        argpdf = (float)mod_twopi(fma(col(kArgpdot), td, col(kArgpo)));
        nodem = (float)mod_twopi(fma(col(kXnodcf), td * td, fma(col(kNodedot), td, col(kNodeo))));
    } else {
        xmdf = fmaf((float)col(kMdot), t, (float)col(kMo));
        argpdf = fmaf((float)col(kArgpdot), t, (float)col(kArgpo));
        nodem = fmaf((float)col(kXnodcf), t2, fmaf((float)col(kNodedot), t, (float)col(kNodeo)));
    }
    float tempa = fmaf(-(float)col(kCc1), t, 1.0f);
    float tempe = (float)col(kBc4) * t;
    float templ = (float)col(kT2cof) * t2;
    float mm = xmdf, argpm = argpdf;
    if (col(kIsimp) == 0.0) {
        const float dm = fmaf((float)col(kEta), cosf(xmdf), 1.0f);
        const float delm = (float)col(kXmcof) * (dm * dm * dm - (float)col(kDelmo));
        const float tho = fmaf((float)col(kOmgcof), t, delm);
        mm = xmdf + tho;
        argpm = argpdf - tho;
        const float t3 = t2 * t, t4 = t3 * t;
        tempa = tempa - (float)col(kD2) * t2 - (float)col(kD3) * t3 - (float)col(kD4) * t4;
        tempe = fmaf((float)col(kBc5), sinf(mm) - (float)col(kSinmao), tempe);
        templ = templ + (float)col(kT3cof) * t3 + t4 * fmaf(t, (float)col(kT5cof), (float)col(kT4cof));
    }
    const float am = (float)col(kAbase) * tempa * tempa;
    const float em = fmaxf((float)col(kEcco) - tempe, 1.0e-6f);
    mm = fmaf((float)col(kNo), templ, mm);

    const float temp = 1.0f / (am * (1.0f - em * em));
    float sa, ca;
    sincosf(argpm, &sa, &ca);
    const float axnl = em * ca;
    const float aynl = fmaf(em, sa, temp * (float)col(kAycof));
    const float u = mm + argpm + temp * (float)col(kXlcof) * axnl;
    float eo1 = u, s = 0.0f, c = 1.0f;
#pragma unroll 1
    for (int it = 0; it < 10; ++it) {
        sincosf(eo1, &s, &c);
        float delta = (u - aynl * c + axnl * s - eo1) / (1.0f - c * axnl - s * aynl);
        delta = fminf(fmaxf(delta, -0.95f), 0.95f);
        eo1 += delta;
        if (fabsf(delta) < 1.0e-6f) break;
    }
    sincosf(eo1, &s, &c);
    const float ecose = axnl * c + aynl * s, esine = axnl * s - aynl * c;
    const float el2 = axnl * axnl + aynl * aynl;
    const float pl = am * (1.0f - el2);
    const float betal = sqrtf(1.0f - el2);
    const float rl = am * (1.0f - ecose);
    const float rdotl = sqrtf(am) * esine / rl;
    const float rvdotl = sqrtf(pl) / rl;
    const float aor = am / rl;
    const float est = esine / (1.0f + betal);
    const float sinu = aor * (s - aynl - axnl * est);
    const float cosu = aor * (c - axnl + aynl * est);
    const float uu = atan2f(sinu, cosu);
    const float sin2u = 2.0f * sinu * cosu, cos2u = 1.0f - 2.0f * sinu * sinu;
    const float temp1 = 0.5f * (float)g.j2 / pl, temp2 = temp1 / pl;
    const float w = 1.0f / (am * sqrtf(am));
    const float con41 = (float)col(kCon41), x1mth2 = (float)col(kX1mth2), x7thm1 = (float)col(kX7thm1);
    const float sinio = (float)col(kSinio), cosio = (float)col(kCosio);
    const float mrt = rl * (1.0f - 1.5f * temp2 * betal * con41) + 0.5f * temp1 * x1mth2 * cos2u;
    const float su = uu - 0.25f * temp2 * x7thm1 * sin2u;
    const float xnode = nodem + 1.5f * temp2 * cosio * sin2u;
    const float xinc = atan2f(sinio, cosio) + 1.5f * temp2 * cosio * sinio * cos2u;
    const float mvt = rdotl - w * temp1 * x1mth2 * sin2u;
    const float rvdot = rvdotl + w * temp1 * (x1mth2 * cos2u + 1.5f * con41);
    float ssu, csu, sn, cn, si, ci;
    sincosf(su, &ssu, &csu);
    sincosf(xnode, &sn, &cn);
    sincosf(xinc, &si, &ci);
    const float xmx = -sn * ci, xmy = cn * ci;
    const float ux = xmx * ssu + cn * csu, uy = xmy * ssu + sn * csu, uz = si * ssu;
    const float vx = xmx * csu - cn * ssu, vy = xmy * csu - sn * ssu, vz = si * csu;
    const float rs = mrt * (float)g.radiusEarthKm, vk = (float)g.vkmpersec;
    o.rx = rs * ux; o.ry = rs * uy; o.rz = rs * uz;
    o.vx = (mvt * ux + rvdot * vx) * vk; o.vy = (mvt * uy + rvdot * vy) * vk; o.vz = (mvt * uz + rvdot * vz) * vk;
    o.mrt = mrt;
}

}  // namespace az
