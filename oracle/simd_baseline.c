/*
 * simd_baseline.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restatement of the reference's *SIMD* CPU path -- the thing the 303 M props/s headline measures:
 *   src/simdMath.zig:29-212      sincosN / modTwoPiN / atan2N / pow15N (8 x f64 lanes)
 *   src/Sgp4Batch.zig:15-157     BatchElements(8), propagateBatchDirect
 *   src/Sgp4.zig:646-750         keplerAndPosVel
 *   src/Sdp4Batch.zig:16-135,199-526       Sdp4BatchElements(8), ResonanceCarryBatch, propagateBatchDirect,
 *                                          computeResonanceAccelBatch, dpperBatch
 *   src/Constellation.zig:101-200          classification into SGP4 / SDP4 batches, origIndices, padding
 *   src/Constellation.zig:327-476,478-528  thread fan-out (SGP4 over time or batch ranges, SDP4 over batch ranges with
 *                                          sequential time for the carry), sgp4Core, unifiedSdp4Range, writeOutput, writeZeros
 * Zig cannot be built in this image, so this port (GCC vector extensions, 8 lanes, per-ISA clones
 * chosen at run time like `oma` does, src/dispatch.zig:21) is the CPU baseline bench.py times.
 * Numerics follow the reference's SIMD path on purpose (including its 1e-7 rad atan2 polynomial):
 * it is validated against the scalar oracle at the reference's own SIMD-vs-scalar tolerance
 * (1e-3 km / 1e-6 km/s, src/Sgp4Batch.zig:180-189) in tests/test_simd_baseline.py.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "astroz_oracle.h"

#define NL 8
typedef double v8d __attribute__((vector_size(64), aligned(64)));
typedef int64_t v8i __attribute__((vector_size(64), aligned(64)));

#define AINL static inline __attribute__((always_inline))
#define SPLAT(x) ((v8d){(x), (x), (x), (x), (x), (x), (x), (x)})
#define ISPLAT(x) ((v8i){(x), (x), (x), (x), (x), (x), (x), (x)})

AINL v8d vsel(v8i m, v8d a, v8d b) { return (v8d)(((v8i)a & m) | ((v8i)b & ~m)); } /* @select */
AINL v8d vfma(v8d a, v8d b, v8d c) {                                                /* @mulAdd */
    v8d r;
    for (int i = 0; i < NL; i++) r[i] = __builtin_fma(a[i], b[i], c[i]);
    return r;
}
AINL v8d vsqrt(v8d a) {
    v8d r;
    for (int i = 0; i < NL; i++) r[i] = __builtin_sqrt(a[i]);
    return r;
}
AINL v8d vfloor(v8d a) {
    v8d r;
    for (int i = 0; i < NL; i++) r[i] = __builtin_floor(a[i]);
    return r;
}
AINL v8d vabs(v8d a) { return (v8d)((v8i)a & ISPLAT(0x7fffffffffffffffLL)); }
AINL v8d vmax(v8d a, v8d b) { return vsel(a > b, a, b); }
AINL v8d vmin(v8d a, v8d b) { return vsel(a < b, a, b); }

typedef struct { v8d s, c; } sc8;

/* src/simdMath.zig:29-97 */
AINL sc8 sincos8(v8d angle) {
    const v8d twoOverPi = SPLAT(2.0 / 3.14159265358979323846);
    const v8d piOver2Hi = SPLAT(1.5707963267948966);
    const v8d piOver2Lo = SPLAT(6.123233995736766e-17);
    const v8d roundMagic = SPLAT(6755399441055744.0);
    v8d kFloat = angle * twoOverPi;
    v8d kRounded = kFloat + roundMagic - roundMagic;
    v8d reduced = vfma(-piOver2Hi, kRounded, angle);
    reduced = vfma(-piOver2Lo, kRounded, reduced);
    v8i k = __builtin_convertvector(kRounded, v8i);
    v8d r2 = reduced * reduced;

    v8d sinP = vfma(SPLAT(1.6058936490373178302326e-10), r2, SPLAT(-2.5052106798274583895303e-08));
    sinP = vfma(sinP, r2, SPLAT(2.7557319210152756118515e-06));
    sinP = vfma(sinP, r2, SPLAT(-1.9841269841201840457725e-04));
    sinP = vfma(sinP, r2, SPLAT(8.3333333333333225058715e-03));
    sinP = vfma(sinP, r2, SPLAT(-1.6666666666666666574148e-01));
    v8d sinReduced = vfma(sinP, r2 * reduced, reduced);

    v8d cosP = vfma(SPLAT(2.0876756987868089233269e-09), r2, SPLAT(-2.7557319223933824788682e-07));
    cosP = vfma(cosP, r2, SPLAT(2.4801587301587286645498e-05));
    cosP = vfma(cosP, r2, SPLAT(-1.3888888888888872762458e-03));
    cosP = vfma(cosP, r2, SPLAT(4.1666666666666665319411e-02));
    cosP = vfma(cosP, r2, SPLAT(-4.9999999999999999999583e-01));
    v8d cosReduced = vfma(cosP, r2, SPLAT(1.0));

    v8i swap = (k & ISPLAT(1)) != ISPLAT(0);
    v8d s = vsel(swap, cosReduced, sinReduced);
    v8d c = vsel(swap, sinReduced, cosReduced);
    v8i sinSign = (k & ISPLAT(2)) << 62;
    v8i cosSign = ((k + ISPLAT(1)) & ISPLAT(2)) << 62;
    sc8 out;
    out.s = (v8d)((v8i)s ^ sinSign);
    out.c = (v8d)((v8i)c ^ cosSign);
    return out;
}

/* src/simdMath.zig:110-122 */
AINL v8d mod2pi8(v8d x) {
    const double twoPi = 2.0 * 3.14159265358979323846;
    v8d n = vfloor(x * SPLAT(1.0 / twoPi));
    v8d r = vfma(SPLAT(-twoPi), n, x);
    return vsel(r < SPLAT(0.0), r + SPLAT(twoPi), r);
}

/* src/simdMath.zig:124-177 (degree-17 polynomial, ~1e-7 rad) */
AINL v8d atan28(v8d y, v8d x) {
    v8d ax = vabs(x), ay = vabs(y);
    v8d mx = vmax(ax, ay), mn = vmin(ax, ay);
    v8d t = mn / vmax(mx, SPLAT(1.0e-30));
    v8d t2 = t * t;
    v8d a = SPLAT(0.0028662257);
    a = vfma(a, t2, SPLAT(-0.0161657367));
    a = vfma(a, t2, SPLAT(0.0429096138));
    a = vfma(a, t2, SPLAT(-0.0752896400));
    a = vfma(a, t2, SPLAT(0.1065626393));
    a = vfma(a, t2, SPLAT(-0.1420889944));
    a = vfma(a, t2, SPLAT(0.1999355085));
    a = vfma(a, t2, SPLAT(-0.3333314528));
    a = vfma(a, t2, SPLAT(1.0));
    a = a * t;
    a = vsel(ay > ax, SPLAT(3.14159265358979323846 / 2.0) - a, a);
    a = vsel(x < SPLAT(0.0), SPLAT(3.14159265358979323846) - a, a);
    a = vsel(y < SPLAT(0.0), -a, a);
    return a;
}

/* src/Sgp4Batch.zig:15-75 -- 40 columns incl. 3 splatted gravity constants, vkmpersec and epochJd */
typedef struct {
    v8d xke, j2, radiusEarthKm;
    v8d ecco, inclo, nodeo, argpo, mo, bstar, noUnkozai;
    v8d sinio, cosio, con41, x1mth2, x7thm1;
    v8d mdot, argpdot, nodedot;
    v8d cc1, cc4, cc5, t2cof, omgcof, xnodcf, xlcof, xmcof, aycof, eta, delmo, sinmao;
    v8d d2, d3, d4, t3cof, t4cof, t5cof;
    v8d aBase, vkmpersec, isimpMask, epochJd;
} batch8;

typedef struct { v8d rx, ry, rz, vx, vy, vz; } pv8;

/* the per-lane arguments of keplerAndPosVel after (am, em, mm, argpm, nodem), src/Sgp4.zig:646-665 */
typedef struct {
    v8d inclo, aycof, xlcof, con41, x1mth2, x7thm1, sinio, cosio, xke, j2, radiusEarthKm, vkmpersec;
} kep8;

/* src/Sgp4.zig:646-750 */
AINL pv8 kepler_posvel8(v8d am, v8d em, v8d mm, v8d argpm, v8d nodem, const kep8 *el) {
    const v8d one = SPLAT(1.0), half = SPLAT(0.5), quarter = SPLAT(0.25), oneHalf = SPLAT(1.5), two = SPLAT(2.0);
    v8d temp = one / (am * (one - em * em));
    sc8 a = sincos8(argpm);
    v8d axnl = em * a.c;
    v8d aynl = em * a.s + temp * el->aycof;
    v8d xl = mm + argpm + nodem + temp * el->xlcof * axnl;
    v8d u = xl - nodem;
    v8d eo1 = u, sineo1 = SPLAT(0.0), coseo1 = one;
    for (int it = 0; it < 10; it++) {
        sc8 sc = sincos8(eo1);
        sineo1 = sc.s;
        coseo1 = sc.c;
        v8d delta = (u - aynl * coseo1 + axnl * sineo1 - eo1) / (one - coseo1 * axnl - sineo1 * aynl);
        eo1 = eo1 + vmax(SPLAT(-0.95), vmin(SPLAT(0.95), delta));
        v8i conv = vabs(delta) < SPLAT(1.0e-12);
        int all = 1;
        for (int i = 0; i < NL; i++) all &= (conv[i] != 0);
        if (all) break;
    }
    v8d ecose = axnl * coseo1 + aynl * sineo1;
    v8d esine = axnl * sineo1 - aynl * coseo1;
    v8d el2 = axnl * axnl + aynl * aynl;
    v8d pl = am * (one - el2);
    v8d betal = vsqrt(one - el2);
    v8d rl = am * (one - ecose);
    v8d rdotl = vsqrt(am) * esine / rl;
    v8d rvdotl = vsqrt(pl) / rl;
    v8d aOverR = am / rl;
    v8d esineTerm = esine / (one + betal);
    v8d sinu = aOverR * (sineo1 - aynl - axnl * esineTerm);
    v8d cosu = aOverR * (coseo1 - axnl + aynl * esineTerm);
    u = atan28(sinu, cosu);
    v8d sin2u = two * sinu * cosu;
    v8d cos2u = one - two * sinu * sinu;

    v8d temp1 = half * el->j2 / pl;
    v8d temp2 = temp1 / pl;
    v8d nm = el->xke / (am * vsqrt(am));
    v8d mrt = rl * (one - oneHalf * temp2 * betal * el->con41) + half * temp1 * el->x1mth2 * cos2u;
    v8d su = u - quarter * temp2 * el->x7thm1 * sin2u;
    v8d xnode = nodem + oneHalf * temp2 * el->cosio * sin2u;
    v8d xinc = el->inclo + oneHalf * temp2 * el->cosio * el->sinio * cos2u;
    v8d mvt = rdotl - nm * temp1 * el->x1mth2 * sin2u / el->xke;
    v8d rvdot = rvdotl + nm * temp1 * (el->x1mth2 * cos2u + oneHalf * el->con41) / el->xke;

    sc8 s = sincos8(su), n = sincos8(xnode), inc = sincos8(xinc);
    v8d xmx = -n.s * inc.c;
    v8d xmy = n.c * inc.c;
    v8d ux = xmx * s.s + n.c * s.c;
    v8d uy = xmy * s.s + n.s * s.c;
    v8d uz = inc.s * s.s;
    v8d vx = xmx * s.c - n.c * s.s;
    v8d vy = xmy * s.c - n.s * s.s;
    v8d vz = inc.s * s.c;
    v8d rScaled = mrt * el->radiusEarthKm;
    pv8 o;
    o.rx = rScaled * ux;
    o.ry = rScaled * uy;
    o.rz = rScaled * uz;
    o.vx = (mvt * ux + rvdot * vx) * el->vkmpersec;
    o.vy = (mvt * uy + rvdot * vy) * el->vkmpersec;
    o.vz = (mvt * uz + rvdot * vz) * el->vkmpersec;
    return o;
}

/* src/Sgp4Batch.zig:113-157 */
AINL pv8 propagate_batch8(const batch8 *el, v8d tsince) {
    const v8d one = SPLAT(1.0), zero = SPLAT(0.0);
    v8d t2 = tsince * tsince, t3 = t2 * tsince, t4 = t3 * tsince;
    v8d tempa = one - el->cc1 * tsince;
    v8d tempe = el->bstar * el->cc4 * tsince;
    v8d templ = el->t2cof * t2;
    v8d xmdf = el->mo + el->mdot * tsince;
    v8d argpdf = el->argpo + el->argpdot * tsince;
    v8d nodem = el->nodeo + el->nodedot * tsince + el->xnodcf * t2;

    v8d delomg = el->omgcof * tsince;
    v8d delmtemp = one + el->eta * sincos8(xmdf).c;
    v8d delmHo = el->xmcof * (delmtemp * delmtemp * delmtemp - el->delmo);
    v8d tempHo = delomg + delmHo;
    v8i hoMask = el->isimpMask == zero;
    v8d mm = xmdf + vsel(hoMask, tempHo, zero);
    v8d argpm = argpdf - vsel(hoMask, tempHo, zero);
    tempa = vsel(hoMask, tempa - el->d2 * t2 - el->d3 * t3 - el->d4 * t4, tempa);
    tempe = vsel(hoMask, tempe + el->bstar * el->cc5 * (sincos8(mm).s - el->sinmao), tempe);
    templ = vsel(hoMask, templ + el->t3cof * t3 + t4 * (el->t4cof + tsince * el->t5cof), templ);

    v8d am = el->aBase * tempa * tempa;
    v8d em = vmax(el->ecco - tempe, SPLAT(1.0e-6));
    mm = mod2pi8(mm + el->noUnkozai * templ);
    nodem = mod2pi8(nodem);
    argpm = mod2pi8(argpm);
    kep8 k = { el->inclo, el->aycof, el->xlcof, el->con41, el->x1mth2, el->x7thm1, el->sinio, el->cosio,
               el->xke, el->j2, el->radiusEarthKm, el->vkmpersec };
    return kepler_posvel8(am, em, mm, argpm, nodem, &k);
}

/* ---------------------------------------------------------------- deep space: src/Sdp4Batch.zig */
/* src/simdMath.zig:180-182, 201-212 */
AINL v8d pow15_8(v8d x) { return x * vsqrt(x); }
AINL v8d pow23_8(v8d x) {
    v8d x2 = x * x;
    v8d y = vsqrt(x);
    const v8d third = SPLAT(1.0 / 3.0), twoThird = SPLAT(2.0 / 3.0);
    for (int i = 0; i < 6; i++) y = twoThird * y + third * x2 / (y * y);
    return y;
}
AINL int vany(v8i m) {
    int a = 0;
    for (int i = 0; i < NL; i++) a |= (m[i] != 0);
    return a;
}

/* src/Sdp4Batch.zig:16-125 -- 77 columns (5 splatted constants, 2 masks) */
typedef struct {
    v8d xke, j2, j3oj2, radiusEarthKm, vkmpersec;
    v8d ecco, inclo, nodeo, argpo, mo, bstar, noUnkozai;
    v8d sinio, cosio, con41, x1mth2, x7thm1;
    v8d mdot, argpdot, nodedot;
    v8d cc1, cc4, t2cof, xnodcf, xlcof, aycof, aBase;
    v8d solar_e2, solar_e3, solar_i2, solar_i3, solar_l2, solar_l3, solar_l4, solar_gh2, solar_gh3, solar_gh4, solar_h2, solar_h3;
    v8d lunar_e2, lunar_e3, lunar_i2, lunar_i3, lunar_l2, lunar_l3, lunar_l4, lunar_gh2, lunar_gh3, lunar_gh4, lunar_h2, lunar_h3;
    v8d zmol, zmos, dedt, didt, dmdt, domdt, dnodt;
    v8d hasResonance, isHalfDay;
    v8d d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433;
    v8d del1, del2, del3;
    v8d xlamo, xfact, gsto;
    v8d epochJd;
} sdp4batch8;

typedef struct { v8d atime, xli, xni; } carry8; /* src/Sdp4Batch.zig:128-135 */

/* src/Sdp4.zig:15-52 */
#define SD_ZNS 1.19459e-5
#define SD_ZES 0.01675
#define SD_ZNL 1.5835218e-4
#define SD_ZEL 0.05490
#define SD_FASX2 0.13130908
#define SD_FASX4 2.8843198
#define SD_FASX6 0.37448087
#define SD_G22 5.7686396
#define SD_G32 0.95240898
#define SD_G44 1.8014998
#define SD_G52 1.0508330
#define SD_G54 4.4108898
#define SD_RPTIM 4.37526908801129966e-3
#define SD_STEPP 720.0
#define SD_STEP2 259200.0

typedef struct { v8d xndt, xnddt, xldot; } accel8;

/* src/Sdp4Batch.zig:347-434: both resonance families evaluated for every lane, merged by mask */
AINL accel8 resonance_accel8(const sdp4batch8 *el, v8d xli, v8d xni, v8d atime) {
    const v8d zero = SPLAT(0.0), two = SPLAT(2.0), three = SPLAT(3.0);
    accel8 o;
    o.xldot = xni + el->xfact;
    v8d xomi = el->argpo + el->argpdot * atime;
    v8d x2omi = xomi + xomi;
    v8d x2li = xli + xli;
    sc8 h1 = sincos8(x2omi + xli - SPLAT(SD_G22));
    sc8 h2 = sincos8(xli - SPLAT(SD_G22));
    sc8 h3 = sincos8(xomi + xli - SPLAT(SD_G32));
    sc8 h4 = sincos8(-xomi + xli - SPLAT(SD_G32));
    sc8 h5 = sincos8(x2omi + x2li - SPLAT(SD_G44));
    sc8 h6 = sincos8(x2li - SPLAT(SD_G44));
    sc8 h7 = sincos8(xomi + xli - SPLAT(SD_G52));
    sc8 h8 = sincos8(-xomi + xli - SPLAT(SD_G52));
    sc8 h9 = sincos8(xomi + x2li - SPLAT(SD_G54));
    sc8 h10 = sincos8(-xomi + x2li - SPLAT(SD_G54));
    v8d xndtHalf = el->d2201 * h1.s + el->d2211 * h2.s + el->d3210 * h3.s + el->d3222 * h4.s + el->d4410 * h5.s +
                   el->d4422 * h6.s + el->d5220 * h7.s + el->d5232 * h8.s + el->d5421 * h9.s + el->d5433 * h10.s;
    v8d xnddtHalf = (el->d2201 * h1.c + el->d2211 * h2.c + el->d3210 * h3.c + el->d3222 * h4.c + el->d5220 * h7.c +
                     el->d5232 * h8.c +
                     two * (el->d4410 * h5.c + el->d4422 * h6.c + el->d5421 * h9.c + el->d5433 * h10.c)) * o.xldot;
    sc8 p1 = sincos8(xli - SPLAT(SD_FASX2));
    sc8 p2 = sincos8(two * (xli - SPLAT(SD_FASX4)));
    sc8 p3 = sincos8(three * (xli - SPLAT(SD_FASX6)));
    v8d xndtGeo = el->del1 * p1.s + el->del2 * p2.s + el->del3 * p3.s;
    v8d xnddtGeo = (el->del1 * p1.c + two * el->del2 * p2.c + three * el->del3 * p3.c) * o.xldot;
    v8i isHalf = el->isHalfDay != zero, hasRes = el->hasResonance != zero;
    o.xndt = vsel(hasRes, vsel(isHalf, xndtHalf, xndtGeo), zero);
    o.xnddt = vsel(hasRes, vsel(isHalf, xnddtHalf, xnddtGeo), zero);
    return o;
}

/* src/Sdp4Batch.zig:436-526 */
AINL void dpper8(const sdp4batch8 *el, v8d tsince, v8d *ep, v8d *inclp, v8d *nodep, v8d *argpp, v8d *mp) {
    const v8d two = SPLAT(2.0), half = SPLAT(0.5), quarter = SPLAT(0.25);
    v8d zm = el->zmos + SPLAT(SD_ZNS) * tsince;
    v8d zf = zm + two * SPLAT(SD_ZES) * sincos8(zm).s;
    v8d sinzf = sincos8(zf).s;
    v8d f2 = half * sinzf * sinzf - quarter;
    v8d f3 = -half * sinzf * sincos8(zf).c;
    v8d ses = el->solar_e2 * f2 + el->solar_e3 * f3;
    v8d sis = el->solar_i2 * f2 + el->solar_i3 * f3;
    v8d sls = el->solar_l2 * f2 + el->solar_l3 * f3 + el->solar_l4 * sinzf;
    v8d sghs = el->solar_gh2 * f2 + el->solar_gh3 * f3 + el->solar_gh4 * sinzf;
    v8d shs = el->solar_h2 * f2 + el->solar_h3 * f3;

    zm = el->zmol + SPLAT(SD_ZNL) * tsince;
    zf = zm + two * SPLAT(SD_ZEL) * sincos8(zm).s;
    sinzf = sincos8(zf).s;
    f2 = half * sinzf * sinzf - quarter;
    f3 = -half * sinzf * sincos8(zf).c;
    v8d sel = el->lunar_e2 * f2 + el->lunar_e3 * f3;
    v8d sil = el->lunar_i2 * f2 + el->lunar_i3 * f3;
    v8d sll = el->lunar_l2 * f2 + el->lunar_l3 * f3 + el->lunar_l4 * sinzf;
    v8d sghl = el->lunar_gh2 * f2 + el->lunar_gh3 * f3 + el->lunar_gh4 * sinzf;
    v8d shl = el->lunar_h2 * f2 + el->lunar_h3 * f3;

    v8d pe = ses + sel, pinc = sis + sil, pl = sls + sll, pgh = sghs + sghl, ph = shs + shl;
    *inclp = *inclp + pinc;
    *ep = *ep + pe;
    v8d sinip = sincos8(*inclp).s, cosip = sincos8(*inclp).c;

    v8d ph_norm = ph / sinip;
    v8d pgh_norm = pgh - cosip * ph_norm;
    v8d argpp_norm = *argpp + pgh_norm;
    v8d nodep_norm = *nodep + ph_norm;
    v8d mp_norm = *mp + pl;

    v8d sinop = sincos8(*nodep).s, cosop = sincos8(*nodep).c;
    v8d alfdp = sinip * sinop, betdp = sinip * cosop;
    v8d dalf = ph * cosop + pinc * cosip * sinop;
    v8d dbet = -ph * sinop + pinc * cosip * cosop;
    alfdp = alfdp + dalf;
    betdp = betdp + dbet;
    v8d nodep_mod = mod2pi8(*nodep);
    v8d xls = *mp + *argpp + cosip * nodep_mod;
    v8d dls = pl + pgh - pinc * nodep_mod * sinip;
    v8d xnoh = nodep_mod;
    v8d nodep_lyd = atan28(alfdp, betdp);
    const v8d pi = SPLAT(3.14159265358979323846), twoPi = SPLAT(2.0 * 3.14159265358979323846);
    v8i need = vabs(xnoh - nodep_lyd) > pi;
    v8d adj = vsel(nodep_lyd < xnoh, twoPi, -twoPi);
    nodep_lyd = vsel(need, nodep_lyd + adj, nodep_lyd);
    v8d mp_lyd = *mp + pl;
    v8d argpp_lyd = xls + dls - mp_lyd - cosip * nodep_lyd;

    v8i useNormal = *inclp >= SPLAT(0.2);
    *argpp = vsel(useNormal, argpp_norm, argpp_lyd);
    *nodep = vsel(useNormal, nodep_norm, nodep_lyd);
    *mp = vsel(useNormal, mp_norm, mp_lyd);
}

/* src/Sdp4Batch.zig:199-343.  Returns a kernel-level code (src/simdKernels.zig:30-37); on error `out` is untouched. */
AINL int propagate_sdp4_batch8(const sdp4batch8 *el, v8d tsince, carry8 *carry, pv8 *out) {
    const v8d one = SPLAT(1.0), zero = SPLAT(0.0), half = SPLAT(0.5), quarter = SPLAT(0.25), two = SPLAT(2.0);
    const v8d three = SPLAT(3.0), five = SPLAT(5.0), seven = SPLAT(7.0);
    v8d t2 = tsince * tsince;
    v8d tempa = one - el->cc1 * tsince;
    v8d tempe = el->bstar * el->cc4 * tsince;
    v8d templ = el->t2cof * t2;
    v8d xmdf = el->mo + el->mdot * tsince;
    v8d argpdf = el->argpo + el->argpdot * tsince;
    v8d nodedf = el->nodeo + el->nodedot * tsince;
    v8d nodem = nodedf + el->xnodcf * t2;

    v8d em = el->ecco + el->dedt * tsince;
    v8d inclm = el->inclo + el->didt * tsince;
    v8d argpm = argpdf + el->domdt * tsince;
    nodem = nodem + el->dnodt * tsince;
    v8d mm = xmdf + el->dmdt * tsince;
    v8d nm = el->noUnkozai;

    v8i hasRes = el->hasResonance != zero;
    if (vany(hasRes)) {
        v8i needsRestart = (carry->atime == zero) | ((tsince * carry->atime) <= zero) | (vabs(tsince) < vabs(carry->atime));
        v8i doRestart = needsRestart & hasRes;
        carry->atime = vsel(doRestart, zero, carry->atime);
        carry->xni = vsel(doRestart, el->noUnkozai, carry->xni);
        carry->xli = vsel(doRestart, el->xlamo, carry->xli);
        const v8d stepp = SPLAT(SD_STEPP), step2 = SPLAT(SD_STEP2);
        v8d delt = vsel(tsince > zero, stepp, -stepp);
        v8i active = hasRes & (vabs(tsince - carry->atime) >= stepp);
        while (vany(active)) {
            accel8 a = resonance_accel8(el, carry->xli, carry->xni, carry->atime);
            v8d nxli = carry->xli + a.xldot * delt + a.xndt * step2;
            v8d nxni = carry->xni + a.xndt * delt + a.xnddt * step2;
            v8d nat = carry->atime + delt;
            carry->xli = vsel(active, nxli, carry->xli);
            carry->xni = vsel(active, nxni, carry->xni);
            carry->atime = vsel(active, nat, carry->atime);
            active = hasRes & (vabs(tsince - carry->atime) >= stepp);
        }
        v8d ft = tsince - carry->atime;
        accel8 a = resonance_accel8(el, carry->xli, carry->xni, carry->atime);
        v8d nm_res = carry->xni + a.xndt * ft + a.xnddt * ft * ft * half;
        v8d xl = carry->xli + a.xldot * ft + a.xndt * ft * ft * half;
        v8d theta = mod2pi8(el->gsto + tsince * SPLAT(SD_RPTIM));
        v8d mm_geo = xl - nodem - argpm + theta;
        v8d mm_half = xl - two * nodem + two * theta;
        v8d mm_res = vsel(el->isHalfDay != zero, mm_half, mm_geo);
        nm = vsel(hasRes, el->noUnkozai + (nm_res - el->noUnkozai), nm);
        mm = vsel(hasRes, mm_res, mm);
    }

    if (vany(nm <= zero)) return AZO_DECAYED;
    v8d am = pow23_8(el->xke / nm) * tempa * tempa;
    nm = el->xke / pow15_8(am);
    em = em - tempe;
    if (vany(em >= one)) return AZO_INVALID_ECC;
    em = vmax(em, SPLAT(1.0e-6));
    if (vany(am < SPLAT(0.95))) return AZO_DECAYED;

    mm = mm + el->noUnkozai * templ;
    v8d xlm = mm + argpm + nodem;
    nodem = mod2pi8(nodem);
    argpm = mod2pi8(argpm);
    mm = mod2pi8(xlm - argpm - nodem);

    dpper8(el, tsince, &em, &inclm, &nodem, &argpm, &mm);

    v8i neg = inclm < zero;
    inclm = vsel(neg, -inclm, inclm);
    nodem = vsel(neg, nodem + SPLAT(3.14159265358979323846), nodem);
    argpm = vsel(neg, argpm - SPLAT(3.14159265358979323846), argpm);
    em = vmax(em, SPLAT(1.0e-6));
    if (vany(em >= one)) return AZO_INVALID_ECC;

    v8d sinip = sincos8(inclm).s, cosip = sincos8(inclm).c;
    v8d cosip2 = cosip * cosip;
    kep8 k;
    k.aycof = -half * el->j3oj2 * sinip;
    v8d denom = cosip + one;
    v8d safe = vsel(vabs(denom) > SPLAT(1.5e-12), denom, SPLAT(1.5e-12));
    k.xlcof = -quarter * el->j3oj2 * sinip * (three + five * cosip) / safe;
    k.x1mth2 = one - cosip2;
    k.con41 = three * cosip2 - one;
    k.x7thm1 = seven * cosip2 - one;
    k.inclo = inclm; k.sinio = sinip; k.cosio = cosip;
    k.xke = el->xke; k.j2 = el->j2; k.radiusEarthKm = el->radiusEarthKm; k.vkmpersec = el->vkmpersec;
    *out = kepler_posvel8(am, em, mm, argpm, nodem, &k);
    return AZO_OK;
}

/* ---------------------------------------------------------------- constellation driver */
typedef struct {
    size_t n;                  /* all satellites = rows of the output block */
    size_t nSgp4, nb;          /* near-earth satellites, batches */
    batch8 *batches;           /* 64-byte aligned */
    double *epochs;            /* near-earth epochs, padded */
    uint32_t *sgp4Orig;        /* padded */
    size_t nSdp4, nbd;         /* deep-space satellites, batches */
    sdp4batch8 *dbatches;
    uint32_t *sdp4Orig;        /* padded */
    carry8 *carries;           /* Constellation.sdp4Carries, :88 */
    double refEpoch;
} simd_const;

typedef struct {
    const simd_const *c;
    const double *tbase, *toff, *jdFull;
    size_t nt, t0, t1, b0, b1;
    double *pos, *vel;
    int layout, deep;
} job;

AINL size_t out_base_simd(int layout, size_t row, size_t t, size_t nt, size_t n) { /* Constellation.zig:46-51 */
    return (layout == 0) ? (row * nt + t) * 3 : (t * n + row) * 3;
}

/* per-ISA clones, resolved at load time (the role of oma.addMultiVersion, build.zig:77-79) */
__attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
static void run_range(const job *j) {
    const simd_const *c = j->c;
    if (j->layout == 0) {                              /* unifiedSgp4Range, Constellation.zig:396-404: batch-major */
        for (size_t b = j->b0; b < j->b1; b++) {
            for (size_t t = j->t0; t < j->t1; t++) {
                v8d ts;
                for (int l = 0; l < NL; l++) ts[l] = j->tbase[t] + j->toff[b * NL + l];
                pv8 r = propagate_batch8(&c->batches[b], ts);
                for (int l = 0; l < NL; l++) {
                    size_t sat = b * NL + l;
                    if (sat >= c->nSgp4) break;
                    size_t ob = out_base_simd(0, c->sgp4Orig[sat], t, j->nt, c->n);
                    j->pos[ob] = r.rx[l]; j->pos[ob + 1] = r.ry[l]; j->pos[ob + 2] = r.rz[l];
                    if (j->vel) { j->vel[ob] = r.vx[l]; j->vel[ob + 1] = r.vy[l]; j->vel[ob + 2] = r.vz[l]; }
                }
            }
        }
        return;
    }
    for (size_t t = j->t0; t < j->t1; t++) {          /* unifiedSgp4Range, Constellation.zig:405-409 */
        for (size_t b = j->b0; b < j->b1; b++) {
            v8d ts;
            for (int l = 0; l < NL; l++) ts[l] = j->tbase[t] + j->toff[b * NL + l]; /* sgp4Core :423-426 */
            pv8 r = propagate_batch8(&c->batches[b], ts);
            for (int l = 0; l < NL; l++) {             /* writeOutput :478-509, TEME */
                size_t sat = b * NL + l;
                if (sat >= c->nSgp4) break;
                size_t ob = out_base_simd(1, c->sgp4Orig[sat], t, j->nt, c->n);
                j->pos[ob] = r.rx[l]; j->pos[ob + 1] = r.ry[l]; j->pos[ob + 2] = r.rz[l];
                if (j->vel) { j->vel[ob] = r.vx[l]; j->vel[ob + 1] = r.vy[l]; j->vel[ob + 2] = r.vz[l]; }
            }
        }
    }
}

/* unifiedSdp4Range, Constellation.zig:448-476: a thread owns a range of deep-space batches and walks the time axis in
 * order, so each batch's resonance carry advances monotonically */
__attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
static void run_range_sdp4(const job *j) {
    const simd_const *c = j->c;
    for (size_t t = 0; t < j->nt; t++) {
        const double jdT = j->jdFull[t];
        for (size_t b = j->b0; b < j->b1; b++) {
            v8d ts;
            for (int l = 0; l < NL; l++) ts[l] = (jdT - c->dbatches[b].epochJd[l]) * 1440.0;
            pv8 r;
            int rc = propagate_sdp4_batch8(&c->dbatches[b], ts, &c->carries[b], &r);
            for (int l = 0; l < NL; l++) {
                size_t sat = b * NL + l;
                if (sat >= c->nSdp4) break;
                size_t ob = out_base_simd(j->layout, c->sdp4Orig[sat], t, j->nt, c->n);
                if (rc != AZO_OK) {                    /* writeZeros :511-528 */
                    j->pos[ob] = j->pos[ob + 1] = j->pos[ob + 2] = 0.0;
                    if (j->vel) j->vel[ob] = j->vel[ob + 1] = j->vel[ob + 2] = 0.0;
                } else {
                    j->pos[ob] = r.rx[l]; j->pos[ob + 1] = r.ry[l]; j->pos[ob + 2] = r.rz[l];
                    if (j->vel) { j->vel[ob] = r.vx[l]; j->vel[ob + 1] = r.vy[l]; j->vel[ob + 2] = r.vz[l]; }
                }
            }
        }
    }
}

static void *thread_main(void *p) {
    const job *j = (const job *)p;
    if (j->deep) run_range_sdp4(j);
    else run_range(j);
    return NULL;
}

void azo_simd_free(void *h) {
    simd_const *c = (simd_const *)h;
    if (!c) return;
    free(c->batches); free(c->epochs); free(c->sgp4Orig);
    free(c->dbatches); free(c->sdp4Orig); free(c->carries);
    free(c);
}

/* Constellation.init (Constellation.zig:101-200): classify, batch by 8, pad the last batch of each class with its last
 * real satellite, keep the original indices. */
void *azo_simd_create(const char *const *l1, const char *const *l2, size_t n, int grav) {
    if (n == 0) return NULL;
    simd_const *c = (simd_const *)calloc(1, sizeof *c);
    azo_sgp4 *sg = (azo_sgp4 *)malloc(n * sizeof(azo_sgp4));
    azo_sdp4 *sd = (azo_sdp4 *)malloc(n * sizeof(azo_sdp4));
    uint32_t *so = (uint32_t *)malloc(n * sizeof(uint32_t)), *dorig = (uint32_t *)malloc(n * sizeof(uint32_t));
    azo_grav g = azo_gravity(grav);
    int bad = 0;
    c->n = n;
    for (size_t i = 0; i < n && !bad; i++) {           /* :115-126 */
        azo_tle t;
        if (azo_tle_parse(l1[i], l2[i], &t) != AZO_OK) { bad = 1; break; }
        int e = azo_sgp4_init(&t, grav, &sg[c->nSgp4]);
        if (e == AZO_OK) { so[c->nSgp4++] = (uint32_t)i; }
        else if (e == AZO_DEEP_SPACE && azo_sdp4_init(&t, grav, &sd[c->nSdp4]) == AZO_OK) { dorig[c->nSdp4++] = (uint32_t)i; }
        else bad = 1;
    }
    if (bad) { free(sg); free(sd); free(so); free(dorig); free(c); return NULL; }
    c->nb = (c->nSgp4 + NL - 1) / NL;
    c->nbd = (c->nSdp4 + NL - 1) / NL;
    if (c->nb) {
        if (posix_memalign((void **)&c->batches, 64, c->nb * sizeof(batch8))) bad = 1;
        c->epochs = (double *)malloc(c->nb * NL * sizeof(double));
        c->sgp4Orig = (uint32_t *)malloc(c->nb * NL * sizeof(uint32_t));
    }
    if (c->nbd && !bad) {
        if (posix_memalign((void **)&c->dbatches, 64, c->nbd * sizeof(sdp4batch8))) bad = 1;
        if (!bad && posix_memalign((void **)&c->carries, 64, c->nbd * sizeof(carry8))) bad = 1;
        c->sdp4Orig = (uint32_t *)malloc(c->nbd * NL * sizeof(uint32_t));
    }
    if (bad) { free(sg); free(sd); free(so); free(dorig); azo_simd_free(c); return NULL; }
    for (size_t s = 0; s < c->nb * NL; s++) {
        size_t src = s < c->nSgp4 ? s : c->nSgp4 - 1;  /* pad with the last real satellite, :146 */
        const azo_sgp4 e = sg[src];
        batch8 *b = &c->batches[s / NL];
        int l = (int)(s % NL);
#define PUT(f) b->f[l] = e.f
        PUT(ecco); PUT(inclo); PUT(nodeo); PUT(argpo); PUT(mo); PUT(bstar); PUT(noUnkozai);
        PUT(sinio); PUT(cosio); PUT(con41); PUT(x1mth2); PUT(x7thm1); PUT(mdot); PUT(argpdot); PUT(nodedot);
        PUT(cc1); PUT(cc4); PUT(cc5); PUT(t2cof); PUT(omgcof); PUT(xnodcf); PUT(xlcof); PUT(xmcof); PUT(aycof);
        PUT(eta); PUT(delmo); PUT(sinmao); PUT(d2); PUT(d3); PUT(d4); PUT(t3cof); PUT(t4cof); PUT(t5cof);
        PUT(aBase); PUT(vkmpersec); PUT(epochJd);
#undef PUT
        b->isimpMask[l] = e.isimp ? 1.0 : 0.0;
        b->xke[l] = g.xke; b->j2[l] = g.j2; b->radiusEarthKm[l] = g.radiusEarthKm;
        c->epochs[s] = e.epochJd;
        c->sgp4Orig[s] = so[src];
    }
    for (size_t s = 0; s < c->nbd * NL; s++) {         /* Sdp4Batch.initFromElements, Sdp4Batch.zig:148-197 */
        size_t src = s < c->nSdp4 ? s : c->nSdp4 - 1;
        const azo_sdp4 e = sd[src];
        sdp4batch8 *b = &c->dbatches[s / NL];
        int l = (int)(s % NL);
#define PUTS(f) b->f[l] = e.s.f
#define PUTD(f) b->f[l] = e.f
        PUTS(ecco); PUTS(inclo); PUTS(nodeo); PUTS(argpo); PUTS(mo); PUTS(bstar); PUTS(noUnkozai);
        PUTS(sinio); PUTS(cosio); PUTS(con41); PUTS(x1mth2); PUTS(x7thm1); PUTS(mdot); PUTS(argpdot); PUTS(nodedot);
        PUTS(cc1); PUTS(cc4); PUTS(t2cof); PUTS(xnodcf); PUTS(xlcof); PUTS(aycof); PUTS(aBase); PUTS(epochJd);
        PUTD(zmol); PUTD(zmos); PUTD(dedt); PUTD(didt); PUTD(dmdt); PUTD(domdt); PUTD(dnodt);
        PUTD(d2201); PUTD(d2211); PUTD(d3210); PUTD(d3222); PUTD(d4410); PUTD(d4422); PUTD(d5220); PUTD(d5232);
        PUTD(d5421); PUTD(d5433); PUTD(del1); PUTD(del2); PUTD(del3); PUTD(xlamo); PUTD(xfact); PUTD(gsto);
#undef PUTS
#undef PUTD
#define PUTP(w, f) b->w##_##f[l] = e.w.f
        PUTP(solar, e2); PUTP(solar, e3); PUTP(solar, i2); PUTP(solar, i3); PUTP(solar, l2); PUTP(solar, l3); PUTP(solar, l4);
        PUTP(solar, gh2); PUTP(solar, gh3); PUTP(solar, gh4); PUTP(solar, h2); PUTP(solar, h3);
        PUTP(lunar, e2); PUTP(lunar, e3); PUTP(lunar, i2); PUTP(lunar, i3); PUTP(lunar, l2); PUTP(lunar, l3); PUTP(lunar, l4);
        PUTP(lunar, gh2); PUTP(lunar, gh3); PUTP(lunar, gh4); PUTP(lunar, h2); PUTP(lunar, h3);
#undef PUTP
        b->hasResonance[l] = e.irez != 0 ? 1.0 : 0.0;
        b->isHalfDay[l] = e.irez == 2 ? 1.0 : 0.0;
        b->xke[l] = g.xke; b->j2[l] = g.j2; b->j3oj2[l] = g.j3oj2; b->radiusEarthKm[l] = g.radiusEarthKm;
        b->vkmpersec[l] = g.xke * g.radiusEarthKm / 60.0;
        c->sdp4Orig[s] = dorig[src];
    }
    c->refEpoch = c->nSgp4 ? c->epochs[0] : 0.0;       /* :139-140 */
    free(sg); free(sd); free(so); free(dorig);
    return c;
}

void azo_simd_counts(void *h, size_t *n, size_t *nSgp4, size_t *nSdp4) {
    simd_const *c = (simd_const *)h;
    if (n) *n = c ? c->n : 0;
    if (nSgp4) *nSgp4 = c ? c->nSgp4 : 0;
    if (nSdp4) *nSdp4 = c ? c->nSdp4 : 0;
}

/* Constellation.propagate (Constellation.zig:245-308) + propagateImpl (:327-385).
 * SGP4 phase: min(work, maxThreads) threads over time ranges (timeMajor) or batch ranges (satelliteMajor).
 * SDP4 phase: threads over batch ranges; sdp4_threads = 0 reproduces the reference's rule -- the deep-space phase gets
 * only the threads the near-earth phase left over, `if (maxThreads > idx) maxThreads - idx else 1` (:361), i.e. ONE
 * thread whenever the near-earth work already used them all; sdp4_threads > 0 gives the phase that many threads
 * (bench.py reports both and keeps the faster as the baseline). */
int azo_simd_propagate2(void *h, const double *jd, const double *fr, size_t nt, double *pos, double *vel, int layout,
                        int nthreads, int sdp4_threads) {
    simd_const *c = (simd_const *)h;
    if (!c || nt == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 128) nthreads = 128;                /* MaxThreads, Constellation.zig:44 */
    double *tbase = (double *)malloc(nt * sizeof(double));
    double *jdFull = (double *)malloc(nt * sizeof(double));
    double *toff = (double *)malloc((c->nb ? c->nb : 1) * NL * sizeof(double));
    const double ref = c->refEpoch;
    for (size_t t = 0; t < nt; t++) {
        jdFull[t] = jd[t] + fr[t];                     /* :266-269 */
        tbase[t] = (jdFull[t] - ref) * 1440.0;
    }
    for (size_t s = 0; s < c->nb * NL; s++) toff[s] = (ref - c->epochs[s]) * 1440.0;
    for (size_t b = 0; b < c->nbd; b++) {              /* a fresh call restarts the integrator when time runs backwards;
                                                          starting from initCarry is what a new Constellation does */
        for (int l = 0; l < NL; l++) {
            c->carries[b].atime[l] = 0.0;
            c->carries[b].xli[l] = c->dbatches[b].xlamo[l];
            c->carries[b].xni[l] = c->dbatches[b].noUnkozai[l];
        }
    }

    pthread_t th[256];
    job jobs[256];
    size_t started = 0, idx = 0;
    if (c->nb) {
        size_t work = (layout == 0) ? c->nb : nt;
        size_t nthr = (size_t)nthreads < work ? (size_t)nthreads : work;
        size_t per = (work + nthr - 1) / nthr;
        for (size_t i = 0; i < nthr; i++) {
            size_t a = i * per, b = a + per < work ? a + per : work;
            if (a >= b) break;
            job *j = &jobs[started];
            memset(j, 0, sizeof *j);
            j->c = c; j->tbase = tbase; j->toff = toff; j->jdFull = jdFull; j->nt = nt; j->pos = pos; j->vel = vel;
            j->layout = layout;
            if (layout == 0) { j->b0 = a; j->b1 = b; j->t0 = 0; j->t1 = nt; }
            else { j->t0 = a; j->t1 = b; j->b0 = 0; j->b1 = c->nb; }
            if (pthread_create(&th[started], NULL, thread_main, j) != 0) run_range(j);
            else started++;
            idx++;
        }
    }
    if (c->nbd) {
        size_t work = c->nbd;
        size_t remaining = sdp4_threads > 0 ? (size_t)sdp4_threads
                                            : ((size_t)nthreads > idx ? (size_t)nthreads - idx : 1);
        size_t nthr = remaining < work ? remaining : work;
        if (nthr > 128) nthr = 128;
        size_t per = (work + nthr - 1) / nthr;
        for (size_t i = 0; i < nthr; i++) {
            size_t a = i * per, b = a + per < work ? a + per : work;
            if (a >= b) break;
            job *j = &jobs[started];
            memset(j, 0, sizeof *j);
            j->c = c; j->jdFull = jdFull; j->nt = nt; j->pos = pos; j->vel = vel; j->layout = layout; j->deep = 1;
            j->b0 = a; j->b1 = b;
            if (pthread_create(&th[started], NULL, thread_main, j) != 0) run_range_sdp4(j);
            else started++;
        }
    }
    for (size_t i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(tbase);
    free(jdFull);
    free(toff);
    return 0;
}

int azo_simd_propagate(void *h, const double *jd, const double *fr, size_t nt, double *pos, double *vel, int layout,
                       int nthreads) {
    return azo_simd_propagate2(h, jd, fr, nt, pos, vel, layout, nthreads, 0);
}

const char *azo_simd_isa(void) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512bw"))
        return "avx512 (x86-64-v4)";
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return "avx2+fma (x86-64-v3)";
    return "sse2 (baseline)";
}
