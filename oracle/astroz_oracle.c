/*
 * astroz_oracle.c -- TEST INFRASTRUCTURE ONLY (see astroz_oracle.h).
 *
 * Scalar fp64 restatement of the reference's libm path for SGP4/SDP4.  Compile with
 * -O2 -ffp-contract=off: Zig's float mode is strict, plain a*b+c is never fused
 * (SURVEY.md Appendix D).  File:line citations are relative to /root/reference.
 */
#include "astroz_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#define AZ_PI 3.14159265358979323846264338327950288
#define AZ_TWOPI (2.0 * AZ_PI)
#define AZ_DEG2RAD (AZ_PI / 180.0)

/* Zig @mod on floats: result takes the sign of the divisor (floored modulo). */
static double mod_floor(double x, double y) {
    double r = fmod(x, y);
    if (r != 0.0 && ((r < 0.0) != (y < 0.0))) r += y;
    return r;
}

/* ------------------------------------------------------------------ constants.zig:41-64 */
azo_grav azo_gravity(int which) {
    azo_grav g;
    if (which == AZO_WGS72) {
        g.radiusEarthKm = 6378.135;
        g.mu = 398600.8;
        g.j2 = 0.001082616;
        g.j3 = -0.00000253881;
        g.j4 = -0.00000165597;
        g.xke = 0.0743669161331734132;
        g.tumin = 13.44683969695931;
        g.j3oj2 = -0.00234506972242078;
    } else {
        g.radiusEarthKm = 6378.137;
        g.mu = 398600.5;
        g.j2 = 0.00108262998905;
        g.j3 = -0.00000253215306;
        g.j4 = -0.00000161098761;
        g.xke = 0.07436685316871385;
        g.tumin = 13.446851082044981;
        g.j3oj2 = -0.00233899967218727;
    }
    return g;
}

/* ------------------------------------------------------------------ Datetime.zig:222-231 */
double azo_year_doy_to_jd(int fullYear, double doy) {
    double y = (double)fullYear;
    double a = floor((14.0 - 1.0) / 12.0);
    double yy = y + 4800.0 - a;
    double mm = 1.0 + 12.0 * a - 3.0;
    double jdJan1 = 1.0 + floor((153.0 * mm + 2.0) / 5.0) + 365.0 * yy + floor(yy / 4.0) - floor(yy / 100.0) +
                    floor(yy / 400.0) - 32045.0;
    return jdJan1 + doy - 1.5;
}

/* ------------------------------------------------------------------ Tle.zig:49-101,277-304 */
static size_t trimmed_len(const char *s) {
    /* std.mem.trim(raw, " \t") then len; leading blanks are skipped by the caller */
    size_t n = strlen(s);
    while (n > 0 && (s[n - 1] == ' ' || s[n - 1] == '\t' || s[n - 1] == '\r' || s[n - 1] == '\n')) n--;
    return n;
}

/* copy line[start:end) with surrounding spaces trimmed (Tle.zig:277-279) */
static int field(const char *line, size_t start, size_t end, char *buf, size_t bufsz) {
    while (start < end && line[start] == ' ') start++;
    while (end > start && line[end - 1] == ' ') end--;
    size_t n = end - start;
    if (n + 1 > bufsz) return -1;
    memcpy(buf, line + start, n);
    buf[n] = 0;
    return (int)n;
}

static int parse_f64(const char *line, size_t s, size_t e, double *out) {
    char buf[40];
    char *endp;
    int n = field(line, s, e, buf, sizeof buf);
    if (n <= 0) return -1;
    *out = strtod(buf, &endp);
    return (*endp == 0) ? 0 : -1;
}

static int parse_i64(const char *line, size_t s, size_t e, long *out) {
    char buf[40];
    char *endp;
    int n = field(line, s, e, buf, sizeof buf);
    if (n <= 0) return -1;
    *out = strtol(buf, &endp, 10);
    return (*endp == 0) ? 0 : -1;
}

int azo_tle_parse(const char *l1, const char *l2, azo_tle *out) {
    while (*l1 == ' ' || *l1 == '\t') l1++;
    while (*l2 == ' ' || *l2 == '\t') l2++;
    if (trimmed_len(l1) < 69 || trimmed_len(l2) < 69) return AZO_BAD_TLE; /* Tle.zig:50 */

    memset(out, 0, sizeof *out);
    double mant;
    long expo, yr;
    if (parse_f64(l1, 53, 59, &mant)) return AZO_BAD_TLE;   /* Tle.zig:69 */
    if (parse_i64(l1, 59, 61, &expo)) return AZO_BAD_TLE;   /* Tle.zig:70 */
    out->bstar = (mant * 1e-5) * pow(10.0, (double)expo);   /* Tle.zig:71 */
    if (parse_i64(l1, 18, 20, &yr)) return AZO_BAD_TLE;
    if (parse_f64(l1, 20, 32, &out->epochDay)) return AZO_BAD_TLE;
    out->epochYear = (int)yr;
    int full = (yr < 57) ? 2000 + (int)yr : 1900 + (int)yr; /* Tle.zig:298-304 */
    out->epochJd = azo_year_doy_to_jd(full, out->epochDay);
    if (parse_f64(l1, 33, 43, &out->ndot)) return AZO_BAD_TLE;

    /* satellite number incl. Alpha-5 (Tle.zig:281-290) */
    {
        char buf[16];
        int n = field(l1, 2, 7, buf, sizeof buf);
        if (n <= 0) return AZO_BAD_TLE;
        if (buf[0] >= 'A' && buf[0] <= 'Z') {
            out->satnum = (uint32_t)(buf[0] - 'A' + 10) * 10000u + (uint32_t)strtoul(buf + 1, NULL, 10);
        } else {
            out->satnum = (uint32_t)strtoul(buf, NULL, 10);
        }
    }

    double e7;
    if (parse_f64(l2, 26, 33, &e7)) return AZO_BAD_TLE;
    out->ecc = e7 / 1e7;                                    /* Tle.zig:78 */
    if (parse_f64(l2, 8, 16, &out->inclDeg)) return AZO_BAD_TLE;
    if (parse_f64(l2, 17, 25, &out->raanDeg)) return AZO_BAD_TLE;
    if (parse_f64(l2, 34, 42, &out->argpDeg)) return AZO_BAD_TLE;
    if (parse_f64(l2, 43, 51, &out->maDeg)) return AZO_BAD_TLE;
    if (parse_f64(l2, 52, 63, &out->nRevDay)) return AZO_BAD_TLE;
    return AZO_OK;
}

/* ------------------------------------------------------------------ Sgp4.zig:192-417 */
typedef struct { double noKozai, ecco, inclo, nodeo, argpo, mo, bstar; } mean_el;
typedef struct { double noUnkozai, a; } recovered;
typedef struct { double sinio, cosio, cosio2, cosio4; } trig_terms;
typedef struct { double con41, con42, x1mth2, x7thm1; } poly_terms;
typedef struct { double mdot, argpdot, nodedot; } sec_rates;
typedef struct { double cc1, cc4, cc5, t2cof, omgcof, xnodcf, xlcof, xmcof, aycof, eta, delmo, sinmao; } drag_co;

static mean_el extract_mean(const azo_tle *t) { /* Sgp4.zig:192-202 */
    mean_el m;
    m.noKozai = t->nRevDay * AZ_TWOPI / 1440.0;
    m.ecco = t->ecc;
    m.inclo = t->inclDeg * AZ_DEG2RAD;
    m.nodeo = t->raanDeg * AZ_DEG2RAD;
    m.argpo = t->argpDeg * AZ_DEG2RAD;
    m.mo = t->maDeg * AZ_DEG2RAD;
    m.bstar = t->bstar;
    return m;
}

static recovered recover_mean_motion(const mean_el *el, const azo_grav *g) { /* Sgp4.zig:206-228 */
    double cosio = cos(el->inclo);
    double theta2 = cosio * cosio;
    double x3thm1 = 3.0 * theta2 - 1.0;
    double eosq = el->ecco * el->ecco;
    double betao2 = 1.0 - eosq;
    double betao = sqrt(betao2);
    double a1 = pow(g->xke / el->noKozai, 2.0 / 3.0);
    double del1 = 0.75 * g->j2 * x3thm1 / (a1 * a1 * betao * betao2);
    double ao = a1 * (1.0 - del1 * (1.0 / 3.0 + del1 * (1.0 + 134.0 / 81.0 * del1)));
    double delo = 0.75 * g->j2 * x3thm1 / (ao * ao * betao * betao2);
    recovered r;
    r.noUnkozai = el->noKozai / (1.0 + delo);
    r.a = pow(g->xke / r.noUnkozai, 2.0 / 3.0);
    return r;
}

static trig_terms trig_of(double inclo) { /* Sgp4.zig:232-237 */
    trig_terms t;
    t.sinio = sin(inclo);
    t.cosio = cos(inclo);
    t.cosio2 = t.cosio * t.cosio;
    t.cosio4 = t.cosio2 * t.cosio2;
    return t;
}

static poly_terms poly_of(const trig_terms *t) { /* Sgp4.zig:241-249 */
    poly_terms p;
    p.con41 = 3.0 * t->cosio2 - 1.0;
    p.con42 = 1.0 - 5.0 * t->cosio2;
    p.x1mth2 = 1.0 - t->cosio2;
    p.x7thm1 = 7.0 * t->cosio2 - 1.0;
    return p;
}

static sec_rates secular_rates(const mean_el *el, const recovered *rec, const trig_terms *tr, const poly_terms *po,
                               const azo_grav *g) { /* Sgp4.zig:253-284 */
    double omeosq = 1.0 - el->ecco * el->ecco;
    double rteosq = sqrt(omeosq);
    double pinvsq = 1.0 / pow(rec->a * omeosq, 2.0);
    double temp1 = 1.5 * g->j2 * pinvsq * rec->noUnkozai;
    double temp2 = 0.5 * temp1 * g->j2 * pinvsq;
    double temp3 = -0.46875 * g->j4 * pinvsq * pinvsq * rec->noUnkozai;
    sec_rates s;
    s.mdot = rec->noUnkozai + 0.5 * temp1 * rteosq * po->con41 +
             0.0625 * temp2 * rteosq * (13.0 - 78.0 * tr->cosio2 + 137.0 * tr->cosio4);
    s.argpdot = -0.5 * temp1 * po->con42 + 0.0625 * temp2 * (7.0 - 114.0 * tr->cosio2 + 395.0 * tr->cosio4) +
                temp3 * (3.0 - 36.0 * tr->cosio2 + 49.0 * tr->cosio4);
    double xhdot1 = -temp1 * tr->cosio;
    s.nodedot = xhdot1 + (0.5 * temp2 * (4.0 - 19.0 * tr->cosio2) + 2.0 * temp3 * (3.0 - 7.0 * tr->cosio2)) * tr->cosio;
    return s;
}

static drag_co drag_coefficients(const mean_el *el, const recovered *rec, const trig_terms *tr, const poly_terms *po,
                                 double perige, const azo_grav *g) { /* Sgp4.zig:301-382 */
    double omeosq = 1.0 - el->ecco * el->ecco;
    double s;
    if (perige < 156.0) s = (perige < 98.0) ? 20.0 : perige - 78.0;
    else s = 78.0;
    double qtemp = (120.0 - s) / g->radiusEarthKm;
    double sfour = s / g->radiusEarthKm + 1.0;
    double qzms24 = qtemp * qtemp * qtemp * qtemp;

    double pinvsq = 1.0 / pow(rec->a * omeosq, 2.0);
    double tsi = 1.0 / (rec->a - sfour);
    double eta = rec->a * el->ecco * tsi;
    double etasq = eta * eta;
    double eeta = el->ecco * eta;
    double psisq = fabs(1.0 - etasq);
    double coef = qzms24 * pow(tsi, 4.0);
    double coef1 = coef / pow(psisq, 3.5);

    double cc2 = coef1 * rec->noUnkozai *
                 (rec->a * (1.0 + 1.5 * etasq + eeta * (4.0 + etasq)) +
                  0.375 * g->j2 * tsi / psisq * po->con41 * (8.0 + 3.0 * etasq * (8.0 + etasq)));
    drag_co d;
    d.cc1 = el->bstar * cc2;
    double cc3 = (el->ecco > 1.0e-4) ? -2.0 * coef * tsi * g->j3oj2 * rec->noUnkozai * tr->sinio / el->ecco : 0.0;
    d.cc4 = 2.0 * rec->noUnkozai * coef1 * rec->a * omeosq *
            (eta * (2.0 + 0.5 * etasq) + el->ecco * (0.5 + 2.0 * etasq) -
             g->j2 * tsi / (rec->a * psisq) *
                 (-3.0 * po->con41 * (1.0 - 2.0 * eeta + etasq * (1.5 - 0.5 * eeta)) +
                  0.75 * po->x1mth2 * (2.0 * etasq - eeta * (1.0 + etasq)) * cos(2.0 * el->argpo)));
    d.cc5 = 2.0 * coef1 * rec->a * omeosq * (1.0 + 2.75 * (etasq + eeta) + eeta * etasq);

    double temp1 = 1.5 * g->j2 * pinvsq * rec->noUnkozai;
    double xhdot1 = -temp1 * tr->cosio;
    d.xnodcf = 3.5 * omeosq * xhdot1 * d.cc1;
    d.t2cof = 1.5 * d.cc1;
    double xlcofNum = -0.25 * g->j3oj2 * tr->sinio * (3.0 + 5.0 * tr->cosio);
    d.xlcof = xlcofNum / ((fabs(tr->cosio + 1.0) > 1.5e-12) ? 1.0 + tr->cosio : 1.5e-12);
    d.aycof = -0.5 * g->j3oj2 * tr->sinio;
    double delmotemp = 1.0 + eta * cos(el->mo);
    d.delmo = delmotemp * delmotemp * delmotemp;
    d.sinmao = sin(el->mo);
    d.xmcof = (el->ecco > 1.0e-4) ? -(2.0 / 3.0) * coef * el->bstar / eeta : 0.0;
    d.omgcof = el->bstar * cc3 * cos(el->argpo);
    d.eta = eta;
    return d;
}

static void fill_common(azo_sgp4 *o, const azo_tle *tle, const azo_grav *g, const mean_el *m, const recovered *rec,
                        const trig_terms *tr, const poly_terms *po, const sec_rates *sr, const drag_co *d) {
    memset(o, 0, sizeof *o);
    o->grav = *g;
    o->epochJd = tle->epochJd;
    o->noKozai = m->noKozai; o->ecco = m->ecco; o->inclo = m->inclo; o->nodeo = m->nodeo;
    o->argpo = m->argpo; o->mo = m->mo; o->bstar = m->bstar;
    o->noUnkozai = rec->noUnkozai; o->a = rec->a;
    o->sinio = tr->sinio; o->cosio = tr->cosio; o->cosio2 = tr->cosio2; o->cosio4 = tr->cosio4;
    o->con41 = po->con41; o->con42 = po->con42; o->x1mth2 = po->x1mth2; o->x7thm1 = po->x7thm1;
    o->mdot = sr->mdot; o->argpdot = sr->argpdot; o->nodedot = sr->nodedot;
    o->cc1 = d->cc1; o->cc4 = d->cc4; o->cc5 = d->cc5; o->t2cof = d->t2cof; o->omgcof = d->omgcof;
    o->xnodcf = d->xnodcf; o->xlcof = d->xlcof; o->xmcof = d->xmcof; o->aycof = d->aycof;
    o->eta = d->eta; o->delmo = d->delmo; o->sinmao = d->sinmao;
    {
        double ratio = g->xke / rec->noUnkozai;   /* Sgp4.zig:173-176 */
        o->aBase = cbrt(ratio * ratio);
    }
    o->vkmpersec = g->xke * g->radiusEarthKm / 60.0;
    o->isimp = 1;
}

int azo_sgp4_init(const azo_tle *tle, int grav, azo_sgp4 *out) { /* Sgp4.zig:108-180 */
    azo_grav g = azo_gravity(grav);
    mean_el m = extract_mean(tle);
    if (m.ecco < 0.0 || m.ecco >= 1.0) return AZO_INVALID_ECC;
    recovered rec = recover_mean_motion(&m, &g);
    double rp = rec.a * (1.0 - m.ecco);
    if (rp < 1.0) return AZO_DECAYED;
    double period = AZ_TWOPI / rec.noUnkozai;
    if (period > 225.0) return AZO_DEEP_SPACE;

    trig_terms tr = trig_of(m.inclo);
    poly_terms po = poly_of(&tr);
    sec_rates sr = secular_rates(&m, &rec, &tr, &po, &g);
    double perige = (rec.a * (1.0 - m.ecco) - 1.0) * g.radiusEarthKm;
    drag_co d = drag_coefficients(&m, &rec, &tr, &po, perige, &g);
    fill_common(out, tle, &g, &m, &rec, &tr, &po, &sr, &d);

    /* computeHigherOrderDrag, Sgp4.zig:394-417 */
    if (perige < 220.0) {
        out->isimp = 1;
    } else {
        double s = 78.0 / g.radiusEarthKm + 1.0;
        double tsi = 1.0 / (rec.a - s);
        double cc1sq = d.cc1 * d.cc1;
        double d2 = 4.0 * rec.a * tsi * cc1sq;
        double temp = d2 * tsi * d.cc1 / 3.0;
        double d3 = (17.0 * rec.a + s) * temp;
        double d4 = 0.5 * temp * rec.a * tsi * (221.0 * rec.a + 31.0 * s) * d.cc1;
        out->d2 = d2; out->d3 = d3; out->d4 = d4;
        out->t3cof = d2 + 2.0 * cc1sq;
        out->t4cof = 0.25 * (3.0 * d3 + d.cc1 * (12.0 * d2 + 10.0 * cc1sq));
        out->t5cof = 0.2 * (3.0 * d4 + 12.0 * d.cc1 * d3 + 6.0 * d2 * d2 + 15.0 * cc1sq * (2.0 * d2 + cc1sq));
        out->isimp = 0;
    }
    return AZO_OK;
}

/* ------------------------------------------------------------------ Sgp4.zig:419-603 (scalar propagate) */
typedef struct { double mm, argpm, nodem, em, a; } sec_state;
typedef struct { double u, r, rdot, rvdot, betal, sin2u, cos2u, nodem, pl; } kep_state;
typedef struct { double r, rdot, rvdot, u, xnode, xinc; } corr_state;

static sec_state update_secular(const azo_sgp4 *el, double tsince) { /* Sgp4.zig:435-477 */
    double t2 = tsince * tsince;
    double tempa = 1.0 - el->cc1 * tsince;
    double tempe = el->bstar * el->cc4 * tsince;
    double templ = el->t2cof * t2;
    double xmdf = el->mo + el->mdot * tsince;
    double argpdf = el->argpo + el->argpdot * tsince;
    double nodedf = el->nodeo + el->nodedot * tsince;
    double argpm = argpdf, mm = xmdf;
    double nodem = nodedf + el->xnodcf * t2;
    if (!el->isimp) {
        double delomg = el->omgcof * tsince;
        double delmtemp = 1.0 + el->eta * cos(xmdf);
        double delm = el->xmcof * (delmtemp * delmtemp * delmtemp - el->delmo);
        double temp = delomg + delm;
        mm = xmdf + temp;
        argpm = argpdf - temp;
        double t3 = t2 * tsince;
        double t4 = t3 * tsince;
        tempa = tempa - el->d2 * t2 - el->d3 * t3 - el->d4 * t4;
        tempe = tempe + el->bstar * el->cc5 * (sin(mm) - el->sinmao);
        templ = templ + el->t3cof * t3 + t4 * (el->t4cof + tsince * el->t5cof);
    }
    sec_state s;
    s.a = el->aBase * tempa * tempa;
    double em = el->ecco - tempe;
    s.em = fmax(em, 1.0e-6);
    mm = mm + el->noUnkozai * templ;
    double xlm = mm + argpm + nodem;
    s.nodem = mod_floor(nodem, AZ_TWOPI);
    s.argpm = mod_floor(argpm, AZ_TWOPI);
    s.mm = mod_floor(xlm - s.argpm - s.nodem, AZ_TWOPI);
    return s;
}

static kep_state solve_kepler(double aycof, double xlcof, const sec_state *sec) { /* Sgp4.zig:495-546 */
    double temp = 1.0 / (sec->a * (1.0 - sec->em * sec->em));
    double axnl = sec->em * cos(sec->argpm);
    double aynl = sec->em * sin(sec->argpm) + temp * aycof;
    double xl = mod_floor(sec->mm + sec->argpm + sec->nodem + temp * xlcof * axnl, AZ_TWOPI);
    double u = mod_floor(xl - sec->nodem, AZ_TWOPI);
    double eo1 = u, sineo1 = 0.0, coseo1 = 1.0, tem5 = 9999.9;
    unsigned ktr = 1;
    while (fabs(tem5) >= 1.0e-12 && ktr <= 10) {
        sineo1 = sin(eo1);
        coseo1 = cos(eo1);
        tem5 = 1.0 - coseo1 * axnl - sineo1 * aynl;
        tem5 = (u - aynl * coseo1 + axnl * sineo1 - eo1) / tem5;
        if (fabs(tem5) >= 0.95) tem5 = (tem5 > 0.0) ? 0.95 : -0.95;
        eo1 = eo1 + tem5;
        ktr += 1;
    }
    double ecose = axnl * coseo1 + aynl * sineo1;
    double esine = axnl * sineo1 - aynl * coseo1;
    double el2 = axnl * axnl + aynl * aynl;
    kep_state k;
    k.pl = sec->a * (1.0 - el2);
    k.betal = sqrt(1.0 - el2);
    double rl = sec->a * (1.0 - ecose);
    k.rdot = sqrt(sec->a) * esine / rl;
    k.rvdot = sqrt(k.pl) / rl;
    double aOverR = sec->a / rl;
    double esineTerm = esine / (1.0 + k.betal);
    double sinu = aOverR * (sineo1 - aynl - axnl * esineTerm);
    double cosu = aOverR * (coseo1 - axnl + aynl * esineTerm);
    k.u = atan2(sinu, cosu);
    k.r = rl;
    k.sin2u = 2.0 * sinu * cosu;
    k.cos2u = 1.0 - 2.0 * sinu * sinu;
    k.nodem = sec->nodem;
    return k;
}

typedef struct { double inclo, sinio, cosio, con41, x1mth2, x7thm1; } incl_terms;

static corr_state short_period(const azo_grav *g, const incl_terms *it, const kep_state *kep, double nm) {
    /* Sgp4.zig:557-571 */
    double temp = 1.0 / kep->pl;
    double temp1 = 0.5 * g->j2 * temp;
    double temp2 = temp1 * temp;
    corr_state c;
    c.r = kep->r * (1.0 - 1.5 * temp2 * kep->betal * it->con41) + 0.5 * temp1 * it->x1mth2 * kep->cos2u;
    c.u = kep->u - 0.25 * temp2 * it->x7thm1 * kep->sin2u;
    c.xnode = kep->nodem + 1.5 * temp2 * it->cosio * kep->sin2u;
    c.xinc = it->inclo + 1.5 * temp2 * it->cosio * it->sinio * kep->cos2u;
    c.rdot = kep->rdot - nm * temp1 * it->x1mth2 * kep->sin2u / g->xke;
    c.rvdot = kep->rvdot + nm * temp1 * (it->x1mth2 * kep->cos2u + 1.5 * it->con41) / g->xke;
    return c;
}

static void pos_vel(const azo_grav *g, double vkmpersec, const corr_state *st, double r[3], double v[3]) {
    /* Sgp4.zig:573-603 */
    double sinsu = sin(st->u), cossu = cos(st->u);
    double snod = sin(st->xnode), cnod = cos(st->xnode);
    double sini = sin(st->xinc), cosi = cos(st->xinc);
    double xmx = -snod * cosi;
    double xmy = cnod * cosi;
    double ux = xmx * sinsu + cnod * cossu;
    double uy = xmy * sinsu + snod * cossu;
    double uz = sini * sinsu;
    double vx = xmx * cossu - cnod * sinsu;
    double vy = xmy * cossu - snod * sinsu;
    double vz = sini * cossu;
    double rScaled = st->r * g->radiusEarthKm;
    r[0] = rScaled * ux; r[1] = rScaled * uy; r[2] = rScaled * uz;
    v[0] = (st->rdot * ux + st->rvdot * vx) * vkmpersec;
    v[1] = (st->rdot * uy + st->rvdot * vy) * vkmpersec;
    v[2] = (st->rdot * uz + st->rvdot * vz) * vkmpersec;
}

void azo_sgp4_propagate(const azo_sgp4 *el, double tsince, double r[3], double v[3]) { /* Sgp4.zig:419-425 */
    sec_state sec = update_secular(el, tsince);
    double nm = el->grav.xke / pow(sec.a, 1.5);
    kep_state kep = solve_kepler(el->aycof, el->xlcof, &sec);
    incl_terms it = { el->inclo, el->sinio, el->cosio, el->con41, el->x1mth2, el->x7thm1 };
    corr_state c = short_period(&el->grav, &it, &kep, nm);
    pos_vel(&el->grav, el->vkmpersec, &c, r, v);
}

/* ------------------------------------------------------------------ Sdp4.zig:15-52 constants */
#define ZES 0.01675
#define ZEL 0.05490
#define C1SS 2.9864797e-6
#define C1L 4.7968065e-7
#define ZSINIS 0.39785416
#define ZCOSIS 0.91744867
#define ZCOSGS 0.1945905
#define ZSINGS (-0.98088458)
#define ZNS 1.19459e-5
#define ZNL 1.5835218e-4
#define Q22 1.7891679e-6
#define Q31 2.1460748e-6
#define Q33 2.2123015e-7
#define ROOT22 1.7891679e-6
#define ROOT32 3.7393792e-7
#define ROOT44 7.3636953e-9
#define ROOT52 1.1428639e-7
#define ROOT54 2.1765803e-9
#define RPTIM 4.37526908801129966e-3
#define FASX2 0.13130908
#define FASX4 2.8843198
#define FASX6 0.37448087
#define G22 5.7686396
#define G32 0.95240898
#define G44 1.8014998
#define G52 1.0508330
#define G54 4.4108898
#define NEAR_EQ 5.2359877e-2
#define STEPP 720.0
#define STEP2 259200.0

double azo_gstime(double jdut1) { /* Sdp4.zig:277-285 */
    double tut1 = (jdut1 - 2451545.0) / 36525.0;
    double temp = -6.2e-6 * tut1 * tut1 * tut1 + 0.093104 * tut1 * tut1 +
                  (876600.0 * 3600.0 + 8640184.812866) * tut1 + 67310.54841;
    temp = mod_floor(temp * AZ_DEG2RAD / 240.0, AZ_TWOPI);
    if (temp < 0.0) temp += AZ_TWOPI;
    return temp;
}

typedef struct {
    azo_perturb solar, lunar;
    double zmol, zmos;
    double snodm, cnodm, sinim, cosim, sinomm, cosomm, emsq, rtemsq, nm, gam;
    double ss1, ss2, ss3, ss4, ss5, ss6, ss7;
    double s1, s2, s3, s4, s5, s6, s7;
    double sz1, sz2, sz3, sz11, sz12, sz13, sz21, sz22, sz23, sz31, sz32, sz33;
    double z1, z2, z3, z11, z12, z13, z21, z22, z23, z31, z32, z33;
} dscom_out;

static azo_perturb perturb_coeffs(double s1v, double s2v, double s3v, double s4v, double s6v, double s7v, double z1t,
                                  double z2t, double z3t, double z11v, double z12v, double z13v, double z21v,
                                  double z22v, double z23v, double z31v, double z32v, double z33v, double emsq,
                                  double ze) { /* Sdp4.zig:69-105 */
    azo_perturb p;
    p.e2 = 2.0 * s1v * s6v;
    p.e3 = 2.0 * s1v * s7v;
    p.i2 = 2.0 * s2v * z12v;
    p.i3 = 2.0 * s2v * (z13v - z11v);
    p.l2 = -2.0 * s3v * z2t;
    p.l3 = -2.0 * s3v * (z3t - z1t);
    p.l4 = -2.0 * s3v * (-21.0 - 9.0 * emsq) * ze;
    p.gh2 = 2.0 * s4v * z32v;
    p.gh3 = 2.0 * s4v * (z33v - z31v);
    p.gh4 = -18.0 * s4v * ze;
    p.h2 = -2.0 * s2v * z22v;
    p.h3 = -2.0 * s2v * (z23v - z21v);
    return p;
}

static dscom_out dscom(const mean_el *el, const recovered *rec, const trig_terms *trig, double day) {
    /* Sdp4.zig:344-499 */
    dscom_out r;
    memset(&r, 0, sizeof r);
    r.nm = rec->noUnkozai;
    r.snodm = sin(el->nodeo);
    r.cnodm = cos(el->nodeo);
    r.sinomm = sin(el->argpo);
    r.cosomm = cos(el->argpo);
    r.sinim = trig->sinio;
    r.cosim = trig->cosio;
    r.emsq = el->ecco * el->ecco;
    r.rtemsq = sqrt(1.0 - r.emsq);

    double xnodce = mod_floor(4.5236020 - 9.2422029e-4 * day, AZ_TWOPI);
    double stem = sin(xnodce), ctem = cos(xnodce);
    double zcosil = 0.91375164 - 0.03568096 * ctem;
    double zsinil = sqrt(1.0 - zcosil * zcosil);
    double zsinhl = 0.089683511 * stem / zsinil;
    double zcoshl = sqrt(1.0 - zsinhl * zsinhl);
    r.gam = 5.8351514 + 0.0019443680 * day;
    double zx = 0.39785416 * stem / zsinil;
    double zy = zcoshl * ctem + 0.91744867 * zsinhl * stem;
    zx = atan2(zx, zy);
    zx += r.gam - xnodce;
    double zcosgl = cos(zx), zsingl = sin(zx);

    double xnoi = 1.0 / r.nm;
    double betasq = 1.0 - r.emsq;

    double zcosg = ZCOSGS, zsing = ZSINGS, zcosi = ZCOSIS, zsini = ZSINIS;
    double zcosh = r.cnodm, zsinh = r.snodm, cc = C1SS;

    for (int lsflg = 1; lsflg <= 2; lsflg++) {
        double a1 = zcosg * zcosh + zsing * zcosi * zsinh;
        double a3 = -zsing * zcosh + zcosg * zcosi * zsinh;
        double a7 = -zcosg * zsinh + zsing * zcosi * zcosh;
        double a8 = zsing * zsini;
        double a9 = zsing * zsinh + zcosg * zcosi * zcosh;
        double a10 = zcosg * zsini;
        double a2 = r.cosim * a7 + r.sinim * a8;
        double a4 = r.cosim * a9 + r.sinim * a10;
        double a5 = -r.sinim * a7 + r.cosim * a8;
        double a6 = -r.sinim * a9 + r.cosim * a10;

        double x1 = a1 * r.cosomm + a2 * r.sinomm;
        double x2 = a3 * r.cosomm + a4 * r.sinomm;
        double x3 = -a1 * r.sinomm + a2 * r.cosomm;
        double x4 = -a3 * r.sinomm + a4 * r.cosomm;
        double x5 = a5 * r.sinomm;
        double x6 = a6 * r.sinomm;
        double x7 = a5 * r.cosomm;
        double x8 = a6 * r.cosomm;

        double z31v = 12.0 * x1 * x1 - 3.0 * x3 * x3;
        double z32v = 24.0 * x1 * x2 - 6.0 * x3 * x4;
        double z33v = 12.0 * x2 * x2 - 3.0 * x4 * x4;
        double z1v = 3.0 * (a1 * a1 + a2 * a2) + z31v * r.emsq;
        double z2v = 6.0 * (a1 * a3 + a2 * a4) + z32v * r.emsq;
        double z3v = 3.0 * (a3 * a3 + a4 * a4) + z33v * r.emsq;
        double z11v = -6.0 * a1 * a5 + r.emsq * (-24.0 * x1 * x7 - 6.0 * x3 * x5);
        double z12v = -6.0 * (a1 * a6 + a3 * a5) + r.emsq * (-24.0 * (x2 * x7 + x1 * x8) - 6.0 * (x3 * x6 + x4 * x5));
        double z13v = -6.0 * a3 * a6 + r.emsq * (-24.0 * x2 * x8 - 6.0 * x4 * x6);
        double z21v = 6.0 * a2 * a5 + r.emsq * (24.0 * x1 * x5 - 6.0 * x3 * x7);
        double z22v = 6.0 * (a4 * a5 + a2 * a6) + r.emsq * (24.0 * (x2 * x5 + x1 * x6) - 6.0 * (x4 * x7 + x3 * x8));
        double z23v = 6.0 * a4 * a6 + r.emsq * (24.0 * x2 * x6 - 6.0 * x4 * x8);

        double z1t = z1v + z1v + betasq * z31v;
        double z2t = z2v + z2v + betasq * z32v;
        double z3t = z3v + z3v + betasq * z33v;

        double s3v = cc * xnoi;
        double s2v = -0.5 * s3v / r.rtemsq;
        double s4v = s3v * r.rtemsq;
        double s1v = -15.0 * el->ecco * s4v;
        double s5v = x1 * x3 + x2 * x4;
        double s6v = x2 * x3 + x1 * x4;
        double s7v = x2 * x4 - x1 * x3;

        if (lsflg == 1) {
            r.ss1 = s1v; r.ss2 = s2v; r.ss3 = s3v; r.ss4 = s4v; r.ss5 = s5v; r.ss6 = s6v; r.ss7 = s7v;
            r.sz1 = z1t; r.sz2 = z2t; r.sz3 = z3t;
            r.sz11 = z11v; r.sz12 = z12v; r.sz13 = z13v;
            r.sz21 = z21v; r.sz22 = z22v; r.sz23 = z23v;
            r.sz31 = z31v; r.sz32 = z32v; r.sz33 = z33v;
            r.solar = perturb_coeffs(s1v, s2v, s3v, s4v, s6v, s7v, z1t, z2t, z3t, z11v, z12v, z13v, z21v, z22v, z23v,
                                     z31v, z32v, z33v, r.emsq, ZES);
            zcosg = zcosgl; zsing = zsingl; zcosi = zcosil; zsini = zsinil;
            zcosh = zcoshl * r.cnodm + zsinhl * r.snodm;
            zsinh = r.snodm * zcoshl - r.cnodm * zsinhl;
            cc = C1L;
        } else {
            r.s1 = s1v; r.s2 = s2v; r.s3 = s3v; r.s4 = s4v; r.s5 = s5v; r.s6 = s6v; r.s7 = s7v;
            r.z1 = z1t; r.z2 = z2t; r.z3 = z3t;
            r.z11 = z11v; r.z12 = z12v; r.z13 = z13v;
            r.z21 = z21v; r.z22 = z22v; r.z23 = z23v;
            r.z31 = z31v; r.z32 = z32v; r.z33 = z33v;
            r.lunar = perturb_coeffs(s1v, s2v, s3v, s4v, s6v, s7v, z1t, z2t, z3t, z11v, z12v, z13v, z21v, z22v, z23v,
                                     z31v, z32v, z33v, r.emsq, ZEL);
        }
    }
    r.zmol = mod_floor(4.7199672 + 0.22997150 * day - r.gam, AZ_TWOPI);
    r.zmos = mod_floor(6.2565837 + 0.017201977 * day, AZ_TWOPI);
    return r;
}

static double poly_eval(double x, const double *c, int n) { /* Sdp4.zig:671-679 */
    double result = 0.0, xn = 1.0;
    for (int i = 0; i < n; i++) {
        result += c[i] * xn;
        xn *= x;
    }
    return result;
}
#define P3(x, a, b, c) poly_eval((x), (const double[]){ a, b, c }, 3)
#define P4(x, a, b, c, d) poly_eval((x), (const double[]){ a, b, c, d }, 4)

static void dsinit(const mean_el *el, const recovered *rec, const trig_terms *trig, const sec_rates *sr,
                   const dscom_out *dc, double gsto, azo_sdp4 *di) { /* Sdp4.zig:525-657 */
    double eosq = el->ecco * el->ecco;
    double cosisq = trig->cosio2;
    double sini2 = trig->sinio * trig->sinio;
    double xpidot = sr->argpdot + sr->nodedot;

    double ses = dc->ss1 * ZNS * dc->ss5;
    double sis = dc->ss2 * ZNS * (dc->sz11 + dc->sz13);
    double sls = -ZNS * dc->ss3 * (dc->sz1 + dc->sz3 - 14.0 - 6.0 * dc->emsq);
    double sghs = dc->ss4 * ZNS * (dc->sz31 + dc->sz33 - 6.0);
    double shs = -ZNS * dc->ss2 * (dc->sz21 + dc->sz23);

    double inclm = el->inclo;
    if (inclm < NEAR_EQ || inclm > AZ_PI - NEAR_EQ) shs = 0.0;
    if (dc->sinim != 0.0) shs = shs / dc->sinim;
    double sgs = sghs - dc->cosim * shs;

    di->dedt = ses + dc->s1 * ZNL * dc->s5;
    di->didt = sis + dc->s2 * ZNL * (dc->z11 + dc->z13);
    di->dmdt = sls - ZNL * dc->s3 * (dc->z1 + dc->z3 - 14.0 - 6.0 * dc->emsq);
    double sghl = dc->s4 * ZNL * (dc->z31 + dc->z33 - 6.0);
    double shll = -ZNL * dc->s2 * (dc->z21 + dc->z23);
    if (inclm < NEAR_EQ || inclm > AZ_PI - NEAR_EQ) shll = 0.0;

    di->domdt = sgs + sghl;
    di->dnodt = shs;
    if (dc->sinim != 0.0) {
        di->domdt -= dc->cosim / dc->sinim * shll;
        di->dnodt += shll / dc->sinim;
    }

    if (rec->noUnkozai >= 0.00826 && rec->noUnkozai <= 0.00924 && el->ecco >= 0.5) di->irez = 2;
    else if (rec->noUnkozai >= 0.0034906585 && rec->noUnkozai <= 0.0052359877) di->irez = 1;
    else di->irez = 0;

    if (di->irez == 1) {
        double g200 = 1.0 + eosq * (-2.5 + 0.8125 * eosq);
        double g310 = 1.0 + 2.0 * eosq;
        double g300 = 1.0 + eosq * (-6.0 + 6.60937 * eosq);
        double f220 = 0.75 * (1.0 + trig->cosio) * (1.0 + trig->cosio);
        double f311 = 0.9375 * sini2 * (1.0 + 3.0 * trig->cosio) - 0.75 * (1.0 + trig->cosio);
        double f330 = 1.0 + trig->cosio;
        f330 = 1.875 * f330 * f330 * f330;
        double aonv = 1.0 / rec->a;
        double temp1g = 3.0 * dc->nm * dc->nm * aonv * aonv;
        di->del2 = 2.0 * temp1g * f220 * g200 * Q22;
        di->del3 = 3.0 * temp1g * f330 * g300 * Q33 * aonv;
        di->del1 = temp1g * f311 * g310 * Q31 * aonv;
        di->xlamo = mod_floor(el->mo + el->nodeo + el->argpo - gsto, AZ_TWOPI);
        di->xfact = sr->mdot + xpidot - RPTIM + di->dmdt + di->domdt + di->dnodt - rec->noUnkozai;
    } else if (di->irez == 2) {
        double e = el->ecco;
        double g201 = -0.306 - (e - 0.64) * 0.440;
        double g211 = (e <= 0.65) ? P3(e, 3.616, -13.2470, 16.2900) : P4(e, -72.099, 331.819, -508.738, 266.724);
        double g310 = (e <= 0.65) ? P4(e, -19.302, 117.3900, -228.4190, 156.591)
                                  : P4(e, -346.844, 1582.851, -2415.925, 1246.113);
        double g322 = (e <= 0.65) ? P4(e, -18.9068, 109.7927, -214.6334, 146.5816)
                                  : P4(e, -342.585, 1554.908, -2366.899, 1215.972);
        double g410 = (e <= 0.65) ? P4(e, -41.122, 242.6940, -471.0940, 313.953)
                                  : P4(e, -1052.797, 4758.686, -7193.992, 3651.957);
        double g422 = (e <= 0.65) ? P4(e, -146.407, 841.8800, -1629.014, 1083.435)
                                  : P4(e, -3581.690, 16178.110, -24462.770, 12422.520);
        double g520;
        if (e <= 0.65) g520 = P4(e, -532.114, 3017.977, -5740.032, 3708.276);
        else if (e > 0.715) g520 = P4(e, -5149.66, 29936.92, -54087.36, 31324.56);
        else g520 = 1464.74 - 4664.75 * e + 3763.64 * e * e;
        double g521 = (e < 0.7) ? P4(e, -822.71072, 4568.6173, -8491.4146, 5337.524)
                                : P4(e, -51752.104, 218913.95, -309468.16, 146349.42);
        double g532 = (e < 0.7) ? P4(e, -853.66600, 4690.2500, -8624.7700, 5341.400)
                                : P4(e, -40023.880, 170470.89, -242699.48, 115605.82);
        double g533 = (e < 0.7) ? P4(e, -919.22770, 4988.6100, -9064.7700, 5542.21)
                                : P4(e, -37995.780, 161616.52, -229838.20, 109377.94);

        double ci = trig->cosio, si = trig->sinio;
        double f220 = 0.75 * (1.0 + 2.0 * ci + cosisq);
        double f221 = 1.5 * sini2;
        double f321 = 1.875 * si * (1.0 - 2.0 * ci - 3.0 * cosisq);
        double f322 = -1.875 * si * (1.0 + 2.0 * ci - 3.0 * cosisq);
        double f441 = 35.0 * sini2 * f220;
        double f442 = 39.3750 * sini2 * sini2;
        double f522 = 9.84375 * si * (sini2 * (1.0 - 2.0 * ci - 5.0 * cosisq) + 0.33333333 * (-2.0 + 4.0 * ci + 6.0 * cosisq));
        double f523 = si * (4.92187512 * sini2 * (-2.0 - 4.0 * ci + 10.0 * cosisq) + 6.56250012 * (1.0 + 2.0 * ci - 3.0 * cosisq));
        double f542 = 29.53125 * si * (2.0 - 8.0 * ci + cosisq * (-12.0 + 8.0 * ci + 10.0 * cosisq));
        double f543 = 29.53125 * si * (-2.0 - 8.0 * ci + cosisq * (12.0 + 8.0 * ci - 10.0 * cosisq));

        double aonv = 1.0 / rec->a;
        double temp1 = 3.0 * dc->nm * dc->nm * aonv * aonv;
        double temp = temp1 * ROOT22;
        di->d2201 = temp * f220 * g201;
        di->d2211 = temp * f221 * g211;
        temp1 = temp1 * aonv;
        temp = temp1 * ROOT32;
        di->d3210 = temp * f321 * g310;
        di->d3222 = temp * f322 * g322;
        temp1 = temp1 * aonv;
        temp = 2.0 * temp1 * ROOT44;
        di->d4410 = temp * f441 * g410;
        di->d4422 = temp * f442 * g422;
        temp1 = temp1 * aonv;
        temp = temp1 * ROOT52;
        di->d5220 = temp * f522 * g520;
        di->d5232 = temp * f523 * g532;
        temp = 2.0 * temp1 * ROOT54;
        di->d5421 = temp * f542 * g521;
        di->d5433 = temp * f543 * g533;

        di->xlamo = mod_floor(el->mo + el->nodeo + el->nodeo - gsto - gsto, AZ_TWOPI);
        di->xfact = sr->mdot + di->dmdt + 2.0 * (sr->nodedot + di->dnodt - RPTIM) - rec->noUnkozai;
    }
}

int azo_sdp4_init(const azo_tle *tle, int grav, azo_sdp4 *out) { /* Sdp4.zig:174-274 */
    azo_grav g = azo_gravity(grav);
    mean_el m = extract_mean(tle);
    if (m.ecco < 0.0 || m.ecco >= 1.0) return AZO_INVALID_ECC;
    recovered rec = recover_mean_motion(&m, &g);
    double rp = rec.a * (1.0 - m.ecco);
    if (rp < 1.0) return AZO_DECAYED;

    trig_terms tr = trig_of(m.inclo);
    poly_terms po = poly_of(&tr);
    sec_rates sr = secular_rates(&m, &rec, &tr, &po, &g);
    double perige = (rec.a * (1.0 - m.ecco) - 1.0) * g.radiusEarthKm;
    drag_co d = drag_coefficients(&m, &rec, &tr, &po, perige, &g);

    memset(out, 0, sizeof *out);
    fill_common(&out->s, tle, &g, &m, &rec, &tr, &po, &sr, &d); /* isimp = 1, d2..t5cof = 0 (Sdp4.zig:231-242) */

    double gsto = azo_gstime(tle->epochJd);
    double day = tle->epochJd - 2415020.0;
    dscom_out dc = dscom(&m, &rec, &tr, day);
    out->solar = dc.solar;
    out->lunar = dc.lunar;
    out->zmol = dc.zmol;
    out->zmos = dc.zmos;
    dsinit(&m, &rec, &tr, &sr, &dc, gsto, out);
    out->gsto = gsto;
    return AZO_OK;
}

/* Sdp4.zig:824-866 */
static void resonance_accel(const azo_sdp4 *el, double xli, double xni, double atime, double *xndt, double *xnddt,
                            double *xldot) {
    *xldot = xni + el->xfact;
    if (el->irez == 2) {
        double xomi = el->s.argpo + el->s.argpdot * atime;
        double x2omi = xomi + xomi;
        double x2li = xli + xli;
        *xndt = el->d2201 * sin(x2omi + xli - G22) + el->d2211 * sin(xli - G22) + el->d3210 * sin(xomi + xli - G32) +
                el->d3222 * sin(-xomi + xli - G32) + el->d4410 * sin(x2omi + x2li - G44) + el->d4422 * sin(x2li - G44) +
                el->d5220 * sin(xomi + xli - G52) + el->d5232 * sin(-xomi + xli - G52) +
                el->d5421 * sin(xomi + x2li - G54) + el->d5433 * sin(-xomi + x2li - G54);
        *xnddt = (el->d2201 * cos(x2omi + xli - G22) + el->d2211 * cos(xli - G22) + el->d3210 * cos(xomi + xli - G32) +
                  el->d3222 * cos(-xomi + xli - G32) + el->d5220 * cos(xomi + xli - G52) +
                  el->d5232 * cos(-xomi + xli - G52) +
                  2.0 * (el->d4410 * cos(x2omi + x2li - G44) + el->d4422 * cos(x2li - G44) +
                         el->d5421 * cos(xomi + x2li - G54) + el->d5433 * cos(-xomi + x2li - G54))) *
                 *xldot;
    } else {
        *xndt = el->del1 * sin(xli - FASX2) + el->del2 * sin(2.0 * (xli - FASX4)) + el->del3 * sin(3.0 * (xli - FASX6));
        *xnddt = (el->del1 * cos(xli - FASX2) + 2.0 * el->del2 * cos(2.0 * (xli - FASX4)) +
                  3.0 * el->del3 * cos(3.0 * (xli - FASX6))) *
                 *xldot;
    }
}

typedef struct { double em, argpm, inclm, mm, nodem, nm, xli, xni, atime; } dspace_state;

static void dspace(const azo_sdp4 *el, double tsince, dspace_state *s) { /* Sdp4.zig:774-820 */
    s->em += el->dedt * tsince;
    s->inclm += el->didt * tsince;
    s->argpm += el->domdt * tsince;
    s->nodem += el->dnodt * tsince;
    s->mm += el->dmdt * tsince;
    if (el->irez == 0) return;

    if (s->atime == 0.0 || tsince * s->atime <= 0.0 || fabs(tsince) < fabs(s->atime)) {
        s->atime = 0.0;
        s->xni = el->s.noUnkozai;
        s->xli = el->xlamo;
    }
    double delt = (tsince > 0.0) ? STEPP : -STEPP;
    double xndt, xnddt, xldot;
    while (fabs(tsince - s->atime) >= STEPP) {
        resonance_accel(el, s->xli, s->xni, s->atime, &xndt, &xnddt, &xldot);
        s->xli += xldot * delt + xndt * STEP2;
        s->xni += xndt * delt + xnddt * STEP2;
        s->atime += delt;
    }
    double ft = tsince - s->atime;
    resonance_accel(el, s->xli, s->xni, s->atime, &xndt, &xnddt, &xldot);
    s->nm = s->xni + xndt * ft + xnddt * ft * ft * 0.5;
    double xl = s->xli + xldot * ft + xndt * ft * ft * 0.5;
    double theta = mod_floor(el->gsto + tsince * RPTIM, AZ_TWOPI);
    if (el->irez != 2) s->mm = xl - s->nodem - s->argpm + theta;
    else s->mm = xl - 2.0 * s->nodem + 2.0 * theta;
    double dndt = s->nm - el->s.noUnkozai;
    s->nm = el->s.noUnkozai + dndt;
}

static void dpper(const azo_sdp4 *el, double tsince, double *ep, double *inclp, double *nodep, double *argpp,
                  double *mp) { /* Sdp4.zig:681-759 */
    double zm = el->zmos + ZNS * tsince;
    double zf = zm + 2.0 * ZES * sin(zm);
    double sinzf = sin(zf);
    double f2 = 0.5 * sinzf * sinzf - 0.25;
    double f3 = -0.5 * sinzf * cos(zf);
    double ses = el->solar.e2 * f2 + el->solar.e3 * f3;
    double sis = el->solar.i2 * f2 + el->solar.i3 * f3;
    double sls = el->solar.l2 * f2 + el->solar.l3 * f3 + el->solar.l4 * sinzf;
    double sghs = el->solar.gh2 * f2 + el->solar.gh3 * f3 + el->solar.gh4 * sinzf;
    double shs = el->solar.h2 * f2 + el->solar.h3 * f3;

    zm = el->zmol + ZNL * tsince;
    zf = zm + 2.0 * ZEL * sin(zm);
    sinzf = sin(zf);
    f2 = 0.5 * sinzf * sinzf - 0.25;
    f3 = -0.5 * sinzf * cos(zf);
    double sel = el->lunar.e2 * f2 + el->lunar.e3 * f3;
    double sil = el->lunar.i2 * f2 + el->lunar.i3 * f3;
    double sll = el->lunar.l2 * f2 + el->lunar.l3 * f3 + el->lunar.l4 * sinzf;
    double sghl = el->lunar.gh2 * f2 + el->lunar.gh3 * f3 + el->lunar.gh4 * sinzf;
    double shl = el->lunar.h2 * f2 + el->lunar.h3 * f3;

    double pe = ses + sel;
    double pinc = sis + sil;
    double pl = sls + sll;
    double pgh = sghs + sghl;
    double ph = shs + shl;

    *inclp += pinc;
    *ep += pe;
    double sinip = sin(*inclp), cosip = cos(*inclp);

    if (*inclp >= 0.2) {
        ph /= sinip;
        pgh -= cosip * ph;
        *argpp += pgh;
        *nodep += ph;
        *mp += pl;
    } else {
        double sinop = sin(*nodep), cosop = cos(*nodep);
        double alfdp = sinip * sinop;
        double betdp = sinip * cosop;
        double dalf = ph * cosop + pinc * cosip * sinop;
        double dbet = -ph * sinop + pinc * cosip * cosop;
        alfdp += dalf;
        betdp += dbet;
        *nodep = mod_floor(*nodep, AZ_TWOPI);
        double xls = *mp + *argpp + cosip * *nodep;
        double dls = pl + pgh - pinc * *nodep * sinip;
        double xnoh = *nodep;
        *nodep = atan2(alfdp, betdp);
        if (fabs(xnoh - *nodep) > AZ_PI) {
            if (*nodep < xnoh) *nodep += AZ_TWOPI;
            else *nodep -= AZ_TWOPI;
        }
        *mp += pl;
        *argpp = xls + dls - *mp - cosip * *nodep;
    }
}

int azo_sdp4_propagate_carry(const azo_sdp4 *el, double tsince, azo_carry *carry, double r[3], double v[3]) {
    /* Sdp4.zig:881-970 */
    const azo_sgp4 *s = &el->s;
    double t2 = tsince * tsince;
    double tempa = 1.0 - s->cc1 * tsince;
    double tempe = s->bstar * s->cc4 * tsince;
    double templ = s->t2cof * t2;
    double xmdf = s->mo + s->mdot * tsince;
    double argpdf = s->argpo + s->argpdot * tsince;
    double nodedf = s->nodeo + s->nodedot * tsince;
    double nodem_init = nodedf + s->xnodcf * t2;

    dspace_state ds = { s->ecco, argpdf, s->inclo, xmdf, nodem_init, s->noUnkozai, carry->xli, carry->xni, carry->atime };
    dspace(el, tsince, &ds);
    carry->atime = ds.atime;
    carry->xli = ds.xli;
    carry->xni = ds.xni;

    double nm = ds.nm;
    if (nm <= 0.0) return AZO_DECAYED;
    double am = pow(s->grav.xke / nm, 2.0 / 3.0) * tempa * tempa;
    nm = s->grav.xke / pow(am, 1.5);
    double em = ds.em - tempe;
    if (em >= 1.0 || em < -0.001) return AZO_INVALID_ECC;
    if (em < 1.0e-6) em = 1.0e-6;
    if (am < 0.95) return AZO_DECAYED;

    double mm = ds.mm + s->noUnkozai * templ;
    double xlm = mm + ds.argpm + ds.nodem;
    double nodem = mod_floor(ds.nodem, AZ_TWOPI);
    double argpm = mod_floor(ds.argpm, AZ_TWOPI);
    mm = mod_floor(xlm - argpm - nodem, AZ_TWOPI);
    double inclm = ds.inclm;

    dpper(el, tsince, &em, &inclm, &nodem, &argpm, &mm);

    if (inclm < 0.0) {
        inclm = -inclm;
        nodem += AZ_PI;
        argpm -= AZ_PI;
    }
    if (em < 1.0e-6) em = 1.0e-6;
    if (em >= 1.0) return AZO_INVALID_ECC;

    double sinip = sin(inclm), cosip = cos(inclm);
    double cosip2 = cosip * cosip;
    double aycof = -0.5 * s->grav.j3oj2 * sinip;
    double xlcofNum = -0.25 * s->grav.j3oj2 * sinip * (3.0 + 5.0 * cosip);
    double xlcof = xlcofNum / ((fabs(cosip + 1.0) > 1.5e-12) ? 1.0 + cosip : 1.5e-12);
    incl_terms it;
    it.inclo = inclm; it.cosio = cosip; it.sinio = sinip;
    it.x1mth2 = 1.0 - cosip2;
    it.con41 = 3.0 * cosip2 - 1.0;
    it.x7thm1 = 7.0 * cosip2 - 1.0;

    sec_state sec = { mm, argpm, nodem, em, am };
    kep_state kep = solve_kepler(aycof, xlcof, &sec);
    corr_state c = short_period(&s->grav, &it, &kep, nm);
    if (c.r < 1.0) return AZO_DECAYED;
    pos_vel(&s->grav, s->vkmpersec, &c, r, v);
    return AZO_OK;
}

int azo_sdp4_propagate(const azo_sdp4 *el, double tsince, double r[3], double v[3]) { /* Sdp4.zig:868-875 */
    azo_carry c = { 0.0, el->xlamo, el->s.noUnkozai };
    return azo_sdp4_propagate_carry(el, tsince, &c, r, v);
}

/* ------------------------------------------------------------------ frames */
double azo_julian_to_gmst(double jd) { /* WorldCoordinateSystem.zig:146-154 */
    double jdJ2000 = jd - 2451545.0;
    double t = jdJ2000 / 36525.0;
    double gmst = 280.46061837 + 360.98564736629 * jdJ2000 + 0.000387933 * t * t - t * t * t / 38710000.0;
    gmst = mod_floor(gmst, 360.0);
    if (gmst < 0) gmst += 360.0;
    return gmst * AZ_DEG2RAD;
}

void azo_eci_to_ecef(const double p[3], double sinG, double cosG, double out[3]) { /* Constellation.zig:54-56 */
    out[0] = p[0] * cosG + p[1] * sinG;
    out[1] = p[1] * cosG - p[0] * sinG;
    out[2] = p[2];
}

void azo_ecef_to_geodetic(const double ecef[3], double lla[3]) { /* WorldCoordinateSystem.zig:98-121 */
    const double wgs84A = 6378.137;                       /* WorldCoordinateSystem.zig:24-26, constants.zig:52-53 */
    const double f = 1.0 / 298.257223563;
    const double e2 = 2.0 * f - f * f;
    double x = ecef[0], y = ecef[1], z = ecef[2];
    double lon = atan2(y, x);
    double p = sqrt(x * x + y * y);
    double lat = atan2(z, p * (1.0 - e2));
    for (int i = 0; i < 10; i++) {
        double latPrev = lat;
        double sinLat = sin(lat);
        double N = wgs84A / sqrt(1.0 - e2 * sinLat * sinLat);
        lat = atan2(z + e2 * N * sinLat, p);
        if (fabs(lat - latPrev) < 1e-12) break;
    }
    double sinLat = sin(lat), cosLat = cos(lat);
    double N = wgs84A / sqrt(1.0 - e2 * sinLat * sinLat);
    lla[0] = lat;
    lla[1] = lon;
    lla[2] = p / cosLat - N;
}

/* ------------------------------------------------------------------ flat exports for ctypes */
void azo_sgp4_export(const azo_sgp4 *e, double *o) {
    int i = 0;
    o[i++] = e->epochJd; o[i++] = e->noKozai; o[i++] = e->ecco; o[i++] = e->inclo; o[i++] = e->nodeo;
    o[i++] = e->argpo; o[i++] = e->mo; o[i++] = e->bstar; o[i++] = e->noUnkozai; o[i++] = e->a;
    o[i++] = e->sinio; o[i++] = e->cosio; o[i++] = e->cosio2; o[i++] = e->cosio4;
    o[i++] = e->con41; o[i++] = e->con42; o[i++] = e->x1mth2; o[i++] = e->x7thm1;
    o[i++] = e->mdot; o[i++] = e->argpdot; o[i++] = e->nodedot;
    o[i++] = e->cc1; o[i++] = e->cc4; o[i++] = e->cc5; o[i++] = e->t2cof; o[i++] = e->omgcof; o[i++] = e->xnodcf;
    o[i++] = e->xlcof; o[i++] = e->xmcof; o[i++] = e->aycof; o[i++] = e->eta; o[i++] = e->delmo; o[i++] = e->sinmao;
    o[i++] = e->d2; o[i++] = e->d3; o[i++] = e->d4; o[i++] = e->t3cof; o[i++] = e->t4cof; o[i++] = e->t5cof;
    o[i++] = e->aBase; o[i++] = e->vkmpersec; o[i++] = (double)e->isimp;
}

void azo_sdp4_export(const azo_sdp4 *e, double *o) {
    azo_sgp4_export(&e->s, o);
    int i = 48;
    const azo_perturb *p = &e->solar;
    for (int k = 0; k < 2; k++) {
        o[i++] = p->e2; o[i++] = p->e3; o[i++] = p->i2; o[i++] = p->i3; o[i++] = p->l2; o[i++] = p->l3; o[i++] = p->l4;
        o[i++] = p->gh2; o[i++] = p->gh3; o[i++] = p->gh4; o[i++] = p->h2; o[i++] = p->h3;
        p = &e->lunar;
    }
    o[i++] = e->zmol; o[i++] = e->zmos; o[i++] = e->dedt; o[i++] = e->didt; o[i++] = e->dmdt; o[i++] = e->domdt;
    o[i++] = e->dnodt; o[i++] = (double)e->irez;
    o[i++] = e->d2201; o[i++] = e->d2211; o[i++] = e->d3210; o[i++] = e->d3222; o[i++] = e->d4410; o[i++] = e->d4422;
    o[i++] = e->d5220; o[i++] = e->d5232; o[i++] = e->d5421; o[i++] = e->d5433;
    o[i++] = e->del1; o[i++] = e->del2; o[i++] = e->del3; o[i++] = e->xlamo; o[i++] = e->xfact; o[i++] = e->gsto;
}

/* ------------------------------------------------------------------ orchestration oracle */
static size_t out_base(int layout, size_t sat, size_t t, size_t nt, size_t ns) { /* Constellation.zig:46-51 */
    return (layout == 0) ? sat * nt * 3 + t * 3 : t * ns * 3 + sat * 3;
}

static void emit(double *pos, double *vel, size_t ob, int mode, const double r[3], const double v[3], double sinG,
                 double cosG) { /* Constellation.zig:478-509 */
    if (mode == 0) {
        memcpy(pos + ob, r, 24);
        if (vel) memcpy(vel + ob, v, 24);
    } else {
        double e[3];
        azo_eci_to_ecef(r, sinG, cosG, e);
        if (mode == 2) azo_ecef_to_geodetic(e, pos + ob);
        else memcpy(pos + ob, e, 24);
        if (vel) azo_eci_to_ecef(v, sinG, cosG, vel + ob);
    }
}

/* Worker state of azo_constellation_propagate_mt: the satellite loop is independent per satellite (SDP4 carries
 * couple only along time within one satellite), so threads take contiguous satellite ranges; the arithmetic per
 * cell is exactly the single-threaded loop's. */
typedef struct {
    const azo_sgp4 *sg; const azo_sdp4 *sd; const int *kind;
    const double *jd, *fr, *gs, *gc;
    size_t n, nt, i0, stride;
    double refEpoch;
    double *pos, *vel;
    int mode, layout;
    uint8_t *err;
} azo_cp_job;

static void *azo_cp_worker(void *arg) {
    const azo_cp_job *j = (const azo_cp_job *)arg;
    for (size_t i = j->i0; i < j->n; i += j->stride) {
        azo_carry carry = { 0.0, 0.0, 0.0 };
        double off = 0.0;
        if (j->kind[i] == 0) off = (j->refEpoch - j->sg[i].epochJd) * 1440.0; /* Constellation.zig:153 */
        else { carry.xli = j->sd[i].xlamo; carry.xni = j->sd[i].s.noUnkozai; }
        for (size_t t = 0; t < j->nt; t++) {
            double jdFull = j->jd[t] + j->fr[t];
            double r[3], v[3];
            int e = AZO_OK;
            if (j->kind[i] == 0) {
                double tsince = (jdFull - j->refEpoch) * 1440.0 + off; /* Constellation.zig:268,425 */
                azo_sgp4_propagate(&j->sg[i], tsince, r, v);
            } else {
                double tsince = (jdFull - j->sd[i].s.epochJd) * 1440.0; /* Constellation.zig:465 */
                e = azo_sdp4_propagate_carry(&j->sd[i], tsince, &carry, r, v);
            }
            size_t ob = out_base(j->layout, i, t, j->nt, j->n);
            if (e != AZO_OK) { /* zero fill, Constellation.zig:511-528 */
                j->pos[ob] = j->pos[ob + 1] = j->pos[ob + 2] = 0.0;
                if (j->vel) j->vel[ob] = j->vel[ob + 1] = j->vel[ob + 2] = 0.0;
            } else {
                emit(j->pos, j->vel, ob, j->mode, r, v, j->gs[t], j->gc[t]);
            }
            if (j->err) j->err[i * j->nt + t] = (uint8_t)e;
        }
    }
    return NULL;
}

int azo_constellation_propagate_mt(const char *const *l1, const char *const *l2, size_t n, int grav, const double *jd,
                                   const double *fr, size_t nt, double *pos, double *vel, int mode, int layout,
                                   uint8_t *err, int *klass, int threads) {
    azo_sgp4 *sg = (azo_sgp4 *)malloc(sizeof(azo_sgp4) * (n ? n : 1));
    azo_sdp4 *sd = (azo_sdp4 *)malloc(sizeof(azo_sdp4) * (n ? n : 1));
    int *kind = (int *)malloc(sizeof(int) * (n ? n : 1));
    double *gs = (double *)malloc(sizeof(double) * (nt ? nt : 1));
    double *gc = (double *)malloc(sizeof(double) * (nt ? nt : 1));
    int rc = 0;
    double refEpoch = 0.0;
    int haveRef = 0;

    for (size_t i = 0; i < n; i++) { /* Constellation.zig:115-126 */
        azo_tle t;
        if (azo_tle_parse(l1[i], l2[i], &t) != AZO_OK) { rc = AZO_BAD_TLE; goto done; }
        int e = azo_sgp4_init(&t, grav, &sg[i]);
        if (e == AZO_OK) {
            kind[i] = 0;
            if (!haveRef) { refEpoch = t.epochJd; haveRef = 1; } /* Constellation.zig:139-140 */
        } else if (e == AZO_DEEP_SPACE) {
            e = azo_sdp4_init(&t, grav, &sd[i]);
            if (e != AZO_OK) { rc = e; goto done; }
            kind[i] = 1 + sd[i].irez;
        } else { rc = e; goto done; }
        if (klass) klass[i] = kind[i];
    }

    for (size_t t = 0; t < nt; t++) { /* Constellation.zig:276-284 */
        if (mode != 0) {
            double g = azo_julian_to_gmst(jd[t] + fr[t]);
            gs[t] = sin(g);
            gc[t] = cos(g);
        } else { gs[t] = 0.0; gc[t] = 0.0; }
    }

    {
        if (threads < 1) threads = 1;
        if (threads > 256) threads = 256;
        if ((size_t)threads > n) threads = n ? (int)n : 1;
        azo_cp_job jobs[256];
        pthread_t tid[256];
        /* satellites are dealt round-robin so a run of deep-space satellites (slower) spreads over the threads */
        for (int k = 0; k < threads; k++) {
            azo_cp_job *j = &jobs[k];
            j->sg = sg; j->sd = sd; j->kind = kind; j->jd = jd; j->fr = fr; j->gs = gs; j->gc = gc;
            j->n = n; j->nt = nt; j->refEpoch = refEpoch; j->pos = pos; j->vel = vel; j->mode = mode;
            j->layout = layout; j->err = err;
            j->i0 = (size_t)k;
            j->stride = (size_t)threads;
        }
        if (threads == 1) {
            azo_cp_worker(&jobs[0]);
        } else {
            for (int k = 0; k < threads; k++) pthread_create(&tid[k], NULL, azo_cp_worker, &jobs[k]);
            for (int k = 0; k < threads; k++) pthread_join(tid[k], NULL);
        }
    }
done:
    free(sg); free(sd); free(kind); free(gs); free(gc);
    return rc;
}

int azo_constellation_propagate(const char *const *l1, const char *const *l2, size_t n, int grav, const double *jd,
                                const double *fr, size_t nt, double *pos, double *vel, int mode, int layout,
                                uint8_t *err, int *klass) {
    return azo_constellation_propagate_mt(l1, l2, n, grav, jd, fr, nt, pos, vel, mode, layout, err, klass, 1);
}

int azo_satrec_array_sgp4(const char *const *l1, const char *const *l2, size_t n, int grav, const double *jd,
                          const double *fr, size_t nt, double *pos, double *vel) {
    if (nt == 0) return 0;
    double referenceJd = jd[0] + fr[0]; /* api.py:300 */
    for (size_t i = 0; i < n; i++) {
        azo_tle t;
        azo_sgp4 el;
        if (azo_tle_parse(l1[i], l2[i], &t) != AZO_OK) return AZO_BAD_TLE;
        int e = azo_sgp4_init(&t, grav, &el);
        if (e != AZO_OK) return e;
        double off = (referenceJd - el.epochJd) * 1440.0; /* api.py:301 */
        for (size_t k = 0; k < nt; k++) {
            double tm = ((jd[k] + fr[k]) - referenceJd) * 1440.0; /* api.py:302 */
            double r[3], v[3];
            azo_sgp4_propagate(&el, tm + off, r, v); /* Constellation.zig:425 */
            memcpy(pos + (i * nt + k) * 3, r, 24);
            if (vel) memcpy(vel + (i * nt + k) * 3, v, 24);
        }
    }
    return 0;
}

int azo_screen_constellation(const char *const *l1, const char *const *l2, size_t n, int grav, const double *times,
                             size_t nt, const double *epoch_offsets, size_t target_idx, double threshold,
                             double reference_jd, double *outMinDists, uint32_t *outMinT) {
    azo_sgp4 *el = (azo_sgp4 *)malloc(sizeof(azo_sgp4) * (n ? n : 1));
    int rc = 0;
    for (size_t i = 0; i < n; i++) {
        azo_tle t;
        if (azo_tle_parse(l1[i], l2[i], &t) != AZO_OK) { rc = AZO_BAD_TLE; goto done; }
        rc = azo_sgp4_init(&t, grav, &el[i]);
        if (rc != AZO_OK) goto done;
    }
    const double thresholdSq = threshold * threshold;            /* Constellation.zig:698 */
    for (size_t i = 0; i < n; i++) { outMinDists[i] = thresholdSq; outMinT[i] = 0; } /* :703-706 */
    for (size_t ti = 0; ti < nt; ti++) {                         /* :720-751 */
        double g = azo_julian_to_gmst(reference_jd + times[ti] / 1440.0);
        double sinG = sin(g), cosG = cos(g);
        double r[3], v[3], tp[3];
        azo_sgp4_propagate(&el[target_idx], times[ti] + epoch_offsets[target_idx], r, v);
        azo_eci_to_ecef(r, sinG, cosG, tp);
        for (size_t s = 0; s < n; s++) {
            if (s == target_idx) continue;
            double p[3];
            azo_sgp4_propagate(&el[s], times[ti] + epoch_offsets[s], r, v);
            azo_eci_to_ecef(r, sinG, cosG, p);
            double dx = tp[0] - p[0], dy = tp[1] - p[1], dz = tp[2] - p[2];
            double distSq = dx * dx + dy * dy + dz * dz;
            if (distSq < outMinDists[s]) { outMinDists[s] = distSq; outMinT[s] = (uint32_t)ti; }
        }
    }
    for (size_t i = 0; i < n; i++) outMinDists[i] = sqrt(outMinDists[i]); /* :753-755 */
done:
    free(el);
    return rc;
}

/* ------------------------------------------------------------------ conjunction.zig:11-149 */
static uint32_t spatial_hash(int32_t cx, int32_t cy, int32_t cz) { /* conjunction.zig:139-148 */
    uint32_t h = (uint32_t)cx;
    h *= 2654435761u;
    h ^= (uint32_t)cy;
    h *= 2654435761u;
    h ^= (uint32_t)cz;
    h *= 2654435761u;
    return h;
}

size_t azo_coarse_screen(const double *positions, size_t num_sats, size_t num_times, double threshold,
                         const uint8_t *valid_mask, uint32_t *out_pairs, uint32_t *out_t, size_t max_results) {
    const double inv_cell = 1.0 / threshold, thr2 = threshold * threshold;
    const size_t TABLE = 1u << 16;
    const uint32_t MASK = (uint32_t)TABLE - 1u, EMPTY = 0xffffffffu;
    size_t count = 0;
    int32_t *cx = (int32_t *)malloc(sizeof(int32_t) * num_sats * 3);
    int32_t *cy = cx + num_sats, *cz = cx + 2 * num_sats;
    uint32_t *hashes = (uint32_t *)malloc(sizeof(uint32_t) * num_sats);
    uint32_t *head = (uint32_t *)malloc(sizeof(uint32_t) * TABLE);
    uint32_t *next = (uint32_t *)malloc(sizeof(uint32_t) * num_sats);
    for (size_t t = 0; t < num_times; t++) {
        memset(head, 0xff, sizeof(uint32_t) * TABLE);
        for (size_t s = 0; s < num_sats; s++) {
            if (valid_mask && valid_mask[s] == 0) { hashes[s] = EMPTY; continue; }
            size_t base = s * num_times * 3 + t * 3;
            double x = positions[base];
            if (!isfinite(x)) { hashes[s] = EMPTY; continue; }
            cx[s] = (int32_t)floor(x * inv_cell);
            cy[s] = (int32_t)floor(positions[base + 1] * inv_cell);
            cz[s] = (int32_t)floor(positions[base + 2] * inv_cell);
            uint32_t h = spatial_hash(cx[s], cy[s], cz[s]) & MASK;
            hashes[s] = h;
            next[s] = head[h];
            head[h] = (uint32_t)s;
        }
        for (size_t s = 0; s < num_sats; s++) {
            if (hashes[s] == EMPTY) continue;
            size_t bs = s * num_times * 3 + t * 3;
            double sx = positions[bs], sy = positions[bs + 1], sz = positions[bs + 2];
            for (int dx = -1; dx <= 1; dx++)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dz = -1; dz <= 1; dz++) {
                        int32_t ncx = cx[s] + dx, ncy = cy[s] + dy, ncz = cz[s] + dz;
                        uint32_t idx = head[spatial_hash(ncx, ncy, ncz) & MASK];
                        while (idx != EMPTY) {
                            uint32_t other = idx;
                            idx = next[idx];
                            if (other <= s) continue;
                            if (cx[other] != ncx || cy[other] != ncy || cz[other] != ncz) continue;
                            size_t bo = (size_t)other * num_times * 3 + t * 3;
                            double ddx = sx - positions[bo], ddy = sy - positions[bo + 1], ddz = sz - positions[bo + 2];
                            if (ddx * ddx + ddy * ddy + ddz * ddz < thr2) {
                                if (count >= max_results) goto done;
                                out_pairs[count * 2] = (uint32_t)s;
                                out_pairs[count * 2 + 1] = other;
                                out_t[count] = (uint32_t)t;
                                count++;
                            }
                        }
                    }
        }
    }
done:
    free(cx); free(hashes); free(head); free(next);
    return count;
}
