/*
 * simd_baseline.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restatement of the reference's *SIMD* CPU path -- the thing the 303 M props/s headline measures:
 *   src/simdMath.zig:29-212      sincosN / modTwoPiN / atan2N / pow15N (8 x f64 lanes)
 *   src/Sgp4Batch.zig:15-157     BatchElements(8), propagateBatchDirect
 *   src/Sgp4.zig:646-750         keplerAndPosVel
 *   src/Constellation.zig:327-434,478-509  thread fan-out over time ranges, sgp4Core, writeOutput
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

/* src/Sgp4.zig:646-750 */
AINL pv8 kepler_posvel8(v8d am, v8d em, v8d mm, v8d argpm, v8d nodem, const batch8 *el) {
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
    return kepler_posvel8(am, em, mm, argpm, nodem, el);
}

/* ---------------------------------------------------------------- constellation driver */
typedef struct {
    size_t n, nb;      /* satellites, batches */
    batch8 *batches;   /* 64-byte aligned */
    double *epochs;    /* padded */
} simd_const;

typedef struct {
    const simd_const *c;
    const double *tbase, *toff;
    size_t nt, t0, t1, b0, b1;
    double *pos, *vel;
    int layout;
} job;

/* per-ISA clones, resolved at load time (the role of oma.addMultiVersion, build.zig:77-79) */
__attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
static void run_range(const job *j) {
    const simd_const *c = j->c;
    for (size_t t = j->t0; t < j->t1; t++) {          /* unifiedSgp4Range, Constellation.zig:405-409 */
        for (size_t b = j->b0; b < j->b1; b++) {
            v8d ts;
            for (int l = 0; l < NL; l++) ts[l] = j->tbase[t] + j->toff[b * NL + l]; /* sgp4Core :423-426 */
            pv8 r = propagate_batch8(&c->batches[b], ts);
            for (int l = 0; l < NL; l++) {             /* writeOutput :478-509, TEME */
                size_t sat = b * NL + l;
                if (sat >= c->n) break;
                size_t ob = (j->layout == 0) ? (sat * j->nt + t) * 3 : (t * c->n + sat) * 3;
                j->pos[ob] = r.rx[l]; j->pos[ob + 1] = r.ry[l]; j->pos[ob + 2] = r.rz[l];
                if (j->vel) { j->vel[ob] = r.vx[l]; j->vel[ob + 1] = r.vy[l]; j->vel[ob + 2] = r.vz[l]; }
            }
        }
    }
}

static void *thread_main(void *p) {
    run_range((const job *)p);
    return NULL;
}

void *azo_simd_create(const char *const *l1, const char *const *l2, size_t n, int grav) {
    if (n == 0) return NULL;
    simd_const *c = (simd_const *)calloc(1, sizeof *c);
    c->n = n;
    c->nb = (n + NL - 1) / NL;
    if (posix_memalign((void **)&c->batches, 64, c->nb * sizeof(batch8))) { free(c); return NULL; }
    c->epochs = (double *)malloc(c->nb * NL * sizeof(double));
    azo_grav g = azo_gravity(grav);
    for (size_t s = 0; s < c->nb * NL; s++) {
        size_t src = s < n ? s : n - 1;                /* pad with the last real satellite, Constellation.zig:146 */
        azo_tle t;
        azo_sgp4 e;
        if (azo_tle_parse(l1[src], l2[src], &t) != AZO_OK || azo_sgp4_init(&t, grav, &e) != AZO_OK) {
            free(c->batches); free(c->epochs); free(c);
            return NULL;
        }
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
    }
    return c;
}

void azo_simd_free(void *h) {
    simd_const *c = (simd_const *)h;
    if (!c) return;
    free(c->batches);
    free(c->epochs);
    free(c);
}

/* Constellation.propagate, near-earth only (Constellation.zig:245-308,327-358): reference epoch = first
 * satellite's epoch; threads over time ranges (timeMajor) or batch ranges (satelliteMajor). */
int azo_simd_propagate(void *h, const double *jd, const double *fr, size_t nt, double *pos, double *vel, int layout,
                       int nthreads) {
    simd_const *c = (simd_const *)h;
    if (!c || nt == 0) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 128) nthreads = 128;                /* MaxThreads, Constellation.zig:44 */
    double *tbase = (double *)malloc(nt * sizeof(double));
    double *toff = (double *)malloc(c->nb * NL * sizeof(double));
    const double ref = c->epochs[0];
    for (size_t t = 0; t < nt; t++) tbase[t] = ((jd[t] + fr[t]) - ref) * 1440.0;
    for (size_t s = 0; s < c->nb * NL; s++) toff[s] = (ref - c->epochs[s]) * 1440.0;

    size_t work = (layout == 0) ? c->nb : nt;
    size_t nthr = (size_t)nthreads < work ? (size_t)nthreads : work;
    size_t per = (work + nthr - 1) / nthr;
    pthread_t th[128];
    job jobs[128];
    size_t started = 0;
    for (size_t i = 0; i < nthr; i++) {
        size_t a = i * per, b = a + per < work ? a + per : work;
        if (a >= b) break;
        job *j = &jobs[started];
        j->c = c; j->tbase = tbase; j->toff = toff; j->nt = nt; j->pos = pos; j->vel = vel; j->layout = layout;
        if (layout == 0) { j->b0 = a; j->b1 = b; j->t0 = 0; j->t1 = nt; }
        else { j->t0 = a; j->t1 = b; j->b0 = 0; j->b1 = c->nb; }
        if (pthread_create(&th[started], NULL, thread_main, j) != 0) run_range(j);
        else started++;
    }
    for (size_t i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(tbase);
    free(toff);
    return 0;
}

const char *azo_simd_isa(void) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
        __builtin_cpu_supports("avx512bw"))
        return "avx512 (x86-64-v4)";
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return "avx2+fma (x86-64-v3)";
    return "sse2 (baseline)";
}
