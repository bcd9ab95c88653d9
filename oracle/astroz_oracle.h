/*
 * astroz_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Scalar CPU restatement (plain C, fp64, libm) of the batch SGP4/SDP4 path of ATTron/astroz.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (astroz_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED -- checked against every golden vector the reference's own tests hold
 * for this path (tests/test_oracle_golden.py; list in SURVEY.md section 8c).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef ASTROZ_ORACLE_H
#define ASTROZ_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kernel-level error codes, src/simdKernels.zig:30-37 */
enum { AZO_OK = 0, AZO_DECAYED = 1, AZO_INVALID_ECC = 2, AZO_DEEP_SPACE = 3, AZO_OOM = 4, AZO_BAD_TLE = 5 };

/* gravity model selector: 0 = WGS84, 1 = WGS72 (src/c_api/sgp4.zig:17-20) */
enum { AZO_WGS84 = 0, AZO_WGS72 = 1 };

typedef struct {
    double radiusEarthKm, mu, j2, j3, j4, xke, tumin, j3oj2;
} azo_grav; /* src/constants.zig:30-64 */

typedef struct {
    uint32_t satnum;
    int epochYear;      /* two-digit */
    double epochDay;
    double epochJd;     /* single f64, src/Datetime.zig:222-231 */
    double ndot, bstar;
    double inclDeg, raanDeg, ecc, argpDeg, maDeg, nRevDay;
} azo_tle; /* src/Tle.zig:8-29 */

typedef struct {
    azo_grav grav;
    double epochJd;
    double noKozai, ecco, inclo, nodeo, argpo, mo, bstar;
    double noUnkozai, a;
    double sinio, cosio, cosio2, cosio4;
    double con41, con42, x1mth2, x7thm1;
    double mdot, argpdot, nodedot;
    double cc1, cc4, cc5, t2cof, omgcof, xnodcf, xlcof, xmcof, aycof, eta, delmo, sinmao;
    double d2, d3, d4, t3cof, t4cof, t5cof;
    double aBase, vkmpersec;
    int isimp;
} azo_sgp4; /* src/Sgp4.zig:33-94 */

typedef struct { double e2, e3, i2, i3, l2, l3, l4, gh2, gh3, gh4, h2, h3; } azo_perturb; /* src/Sdp4.zig:54-67 */

typedef struct {
    azo_sgp4 s;
    azo_perturb solar, lunar;
    double zmol, zmos, dedt, didt, dmdt, domdt, dnodt;
    int irez;
    double d2201, d2211, d3210, d3222, d4410, d4422, d5220, d5232, d5421, d5433;
    double del1, del2, del3;
    double xlamo, xfact, gsto;
} azo_sdp4; /* src/Sdp4.zig:109-148 */

typedef struct { double atime, xli, xni; } azo_carry; /* src/Sdp4.zig:162-166 */

azo_grav azo_gravity(int which);

/* src/Tle.zig:49-101 ; returns AZO_OK or AZO_BAD_TLE */
int azo_tle_parse(const char *line1, const char *line2, azo_tle *out);
double azo_year_doy_to_jd(int fullYear, double doy);           /* src/Datetime.zig:222-231 */

int azo_sgp4_init(const azo_tle *tle, int grav, azo_sgp4 *out);   /* src/Sgp4.zig:108-180 */
void azo_sgp4_propagate(const azo_sgp4 *el, double tsince, double r[3], double v[3]); /* src/Sgp4.zig:419-603 */

int azo_sdp4_init(const azo_tle *tle, int grav, azo_sdp4 *out);   /* src/Sdp4.zig:174-274 */
int azo_sdp4_propagate(const azo_sdp4 *el, double tsince, double r[3], double v[3]); /* src/Sdp4.zig:868-875 */
int azo_sdp4_propagate_carry(const azo_sdp4 *el, double tsince, azo_carry *c, double r[3], double v[3]); /* :881-970 */

double azo_gstime(double jdut1);                                  /* src/Sdp4.zig:277-285 */
double azo_julian_to_gmst(double jd);                             /* src/WorldCoordinateSystem.zig:146-154 */
void azo_eci_to_ecef(const double p[3], double sinG, double cosG, double out[3]); /* src/Constellation.zig:54-56 */
void azo_ecef_to_geodetic(const double ecef[3], double lla[3]);   /* src/WorldCoordinateSystem.zig:98-121 */

/* flat exports for ctypes: fills out[] in the documented order (see oracle/oracle.py) */
void azo_sgp4_export(const azo_sgp4 *el, double *out /*>=48*/);
void azo_sdp4_export(const azo_sdp4 *el, double *out /*>=112*/);

/*
 * Whole-constellation scalar oracle with the reference's orchestration semantics
 * (src/Constellation.zig:101-200 classify, :245-308 time model, :478-528 output modes / zero fill).
 *   lines1/lines2 : n NUL-terminated TLE lines
 *   mode   : 0 teme, 1 ecef, 2 geodetic ; layout : 0 satelliteMajor, 1 timeMajor
 *   pos/vel: n*nt*3 doubles (vel may be NULL)
 *   err    : n*nt bytes (may be NULL) kernel-level code per cell (0 = ok), satellite-major
 *   klass  : n ints (may be NULL): 0 = SGP4, 1 = SDP4 irez0, 2 = irez1, 3 = irez2
 * Returns 0, or the init error of the first TLE that is neither SGP4 nor SDP4 (Constellation.zig:124).
 */
int azo_constellation_propagate(const char *const *lines1, const char *const *lines2, size_t n, int grav,
                                const double *jd, const double *fr, size_t nt,
                                double *pos, double *vel, int mode, int layout,
                                uint8_t *err, int *klass);
/* Same cells, same arithmetic, satellites dealt to `threads` pthreads (full-size parity runs in seconds). */
int azo_constellation_propagate_mt(const char *const *lines1, const char *const *lines2, size_t n, int grav,
                                   const double *jd, const double *fr, size_t nt,
                                   double *pos, double *vel, int mode, int layout,
                                   uint8_t *err, int *klass, int threads);

/* SatrecArray.sgp4 time model (bindings/python/astroz/api.py:300-302): reference = jd[0]+fr[0]. SGP4 only. */
int azo_satrec_array_sgp4(const char *const *lines1, const char *const *lines2, size_t n, int grav,
                          const double *jd, const double *fr, size_t nt,
                          double *pos_satmajor, double *vel_satmajor);

/* Constellation.screenConstellation (src/Constellation.zig:683-756): near-earth satellites only,
 * tsince = times[t] + epoch_offsets[sat]; ECEF rotation by GMST(reference_jd + t/1440) as the reference. */
int azo_screen_constellation(const char *const *lines1, const char *const *lines2, size_t n, int grav,
                             const double *times, size_t nt, const double *epoch_offsets, size_t target_idx,
                             double threshold, double reference_jd, double *out_min_dists, uint32_t *out_min_t);

/* coarseScreen (bindings/python/src/conjunction.zig:11-149): satellite-major positions [num_sats][num_times][3],
 * cell-list spatial hash per time step.  Returns the number of results written (<= max_results). */
size_t azo_coarse_screen(const double *positions, size_t num_sats, size_t num_times, double threshold,
                         const uint8_t *valid_mask, uint32_t *out_pairs, uint32_t *out_t_indices, size_t max_results);

#ifdef __cplusplus
}
#endif
#endif
