/* oracle/oracle_sc.c -- TEST INFRASTRUCTURE (see oracle_common.h).  Oriented path: mosaic
 * single-crystal Bragg diffraction (SCBragg / GaussMos / GaussOnSphere) and the oriented
 * ProcComposition entry points.  The list of contributing normals ("xs_commul"/"scatcache" of
 * SCBragg::pimpl::Cache, ref: src/scbragg/NCSCBragg.cc:68-79) is materialised here exactly as the
 * reference does, one neutron at a time. */
#include "oracle_common.h"

#define SC_PIHALF 1.5707963267948966192313216916397514420985847
#define SC_2PI 6.2831853071795864769252867665590057683943388
#define SC_ARCSEC 0.00000484813681109535993589914102357947975956353302

static double vdot(orc_vec a, orc_vec b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
static double vmag2(orc_vec a) { return a.x*a.x + a.y*a.y + a.z*a.z; }
static orc_vec vcross(orc_vec a, orc_vec o) { orc_vec r = { a.y*o.z - a.z*o.y, a.z*o.x - a.x*o.z, a.x*o.y - a.y*o.x }; return r; }
/* Vector::normalise, ref: include/NCrystal/internal/utils/NCVector.hh:197-209 */
static orc_vec vnorm(orc_vec v)
{
  double m2 = vmag2(v);
  if (m2 >= 1.0 - 2.0*DBL_EPSILON && m2 <= 1.0 + 2.0*DBL_EPSILON) return v;
  double ff = 1.0/sqrt(m2);
  v.x *= ff; v.y *= ff; v.z *= ff;
  return v;
}

/* sincos_mpi2pi2 / sincos_mpi8pi8 / sincos_0pi32 / cos_mpipi, ref: src/utils/NCMath.cc:96-175, NCMath.hh:368-373 */
static void sincos_mpi2pi2(double A, double* c, double* s)
{
  double x = 0.5*A, m = -x*x;
  double s2 = x*(1.0 + m*(1.66666666666666666666666666666666666666666667e-1 + m*(8.33333333333333333333333333333333333333333333e-3
            + m*(1.98412698412698412698412698412698412698412698e-4 + m*(2.75573192239858906525573192239858906525573192e-6
            + m*(2.50521083854417187750521083854417187750521084e-8 + m*(1.60590438368216145993923771701549479327257105e-10
            + m*(7.64716373181981647590113198578807044415510024e-13))))))));
  double c2m1 = m*(0.5 + m*(4.16666666666666666666666666666666666666666667e-2 + m*(1.38888888888888888888888888888888888888888889e-3
              + m*(2.48015873015873015873015873015873015873015873e-5 + m*(2.75573192239858906525573192239858906525573192e-7
              + m*(2.08767569878680989792100903212014323125434237e-9 + m*(1.14707455977297247138516979786821056662326504e-11
              + m*(4.77947733238738529743820749111754402759693765e-14))))))));
  double k = 2.0*c2m1;
  *s = (k + 2.0)*s2;
  *c = 1.0 + k*(c2m1 + 2.0);
}
static void sincos_mpi8pi8(double A, double* c, double* s)
{
  double x = 0.5*A, m = -x*x;
  double s2 = x*(1.0 + m*(1.66666666666666666666666666666666666666666667e-1 + m*(8.33333333333333333333333333333333333333333333e-3
            + m*(1.98412698412698412698412698412698412698412698e-4 + m*(2.75573192239858906525573192239858906525573192e-6
            + m*(2.50521083854417187750521083854417187750521084e-8))))));
  double c2m1 = m*(0.5 + m*(4.16666666666666666666666666666666666666666667e-2 + m*(1.38888888888888888888888888888888888888888889e-3
              + m*(2.48015873015873015873015873015873015873015873e-5 + m*(2.75573192239858906525573192239858906525573192e-7)))));
  double k = 2.0*c2m1;
  *s = (k + 2.0)*s2;
  *c = 1.0 + k*(c2m1 + 2.0);
}
static void sincos_0pi32(double A, double* c, double* s)
{
  sincos_mpi2pi2(orc_min(A, ORC_PI - A), c, s);
  *c = copysign(*c, SC_PIHALF - A);
}
static double cos_mpipi(double A)
{
  double Aabs = fabs(A), x = orc_min(Aabs, ORC_PI - Aabs), m = -x*x;
  double c = 1.0 + m*(0.5 + m*(4.16666666666666666666666666666666666666666667e-2 + m*(1.38888888888888888888888888888888888888888889e-3
           + m*(2.48015873015873015873015873015873015873015873e-5 + m*(2.75573192239858906525573192239858906525573192e-7
           + m*(2.08767569878680989792100903212014323125434237e-9 + m*(1.14707455977297247138516979786821056662326504e-11
           + m*(4.77947733238738529743820749111754402759693765e-14 + m*(1.56192069685862264622163643500573334235194041e-16
           + m*(4.1103176233121648584779906184361403746103695e-19 + m*(8.89679139245057328674889744250246834331248809e-22)))))))))));
  return copysign(c, SC_PIHALF - Aabs);
}

/* CubicSpline::evalUnbounded via SplinedLookupTable::eval, ref: include/NCrystal/internal/utils/NCSpline.hh:122-135,158-160 */
static double lut_eval(const orc_lut* L, double xin)
{
  double x = (xin - L->a)*L->invdelta;
  long long ll = (long long)x;
  if (ll < 0) ll = 0;
  int idx = ll < (long long)L->nm2 ? (int)ll : L->nm2;
  double b = x - idx, a = 1.0 - b;
  const double* it = L->data + 2*idx;
  double tmp = a*it[0];
  double tmp2 = (a*a*a - a)*it[1];
  tmp += b*it[2];
  tmp2 += (b*b*b - b)*it[3];
  return tmp + 0.166666666666666666666666666666666666666666666666666667*tmp2;
}
/* GaussOnSphere::evalCosXInRange / evalCosX, ref: include/NCrystal/internal/phys_utils/NCGaussOnSphere.hh:165-181 */
static double gos_cosx_inrange(const orc_sc* S, double cx) { return orc_max(0.0, lut_eval(&S->evalcosx, cx)); }
static double gos_cosx(const orc_sc* S, double cx) { return cx >= S->evalcosx.a ? gos_cosx_inrange(S, cx) : 0.0; }

/* CosSinGridGen, ref: include/NCrystal/internal/utils/NCMath.hh:164-190,489-525 */
typedef struct { double c, s, cd, sd, phimax, negdelta; unsigned left, recalc; } csgrid;
static void csgrid_init(csgrid* g, unsigned n, double offset, double delta)
{
  g->left = n-1; g->recalc = ((127u + (n/128u)*128u) - n);
  g->phimax = offset + (n-1)*delta; g->negdelta = -delta;
  sincos_0pi32(offset, &g->c, &g->s);
  sincos_mpi8pi8(delta, &g->cd, &g->sd);
}
static int csgrid_step(csgrid* g)
{
  if (!g->left) return 0;
  --g->left;
  if ((g->left + g->recalc) % 128u) {
    double c = g->c*g->cd - g->s*g->sd;
    g->s = g->c*g->sd + g->s*g->cd;
    g->c = c;
  } else {
    double v = g->phimax + g->negdelta*g->left;
    g->c = cos(v); g->s = sin(v);
  }
  return 1;
}
/* GOSCircleInt::evalFuncMany / evalFuncManySum / accept, ref: src/phys_utils/NCGaussOnSphere.cc:83-141 */
static void gos_many(const orc_sc* S, double sasg, double cacg, double* f, unsigned n, double off, double d)
{
  csgrid g; csgrid_init(&g, n, off, d);
  unsigned i = 0;
  do { f[i++] = gos_cosx_inrange(S, sasg*g.c + cacg); } while (csgrid_step(&g));
}
static double gos_manysum(const orc_sc* S, double sasg, double cacg, unsigned n, double off, double d)
{
  csgrid g; csgrid_init(&g, n, off, d);
  double sum = 0.;
  do { sum += gos_cosx_inrange(S, sasg*g.c + cacg); } while (csgrid_step(&g));
  return sum;
}
static int gos_accept(double acc, unsigned level, double prev, double est)
{
  if (fabs(prev - est) <= acc*fabs(est)) return 1;
  if (level < 11) return 0;
  return 1;
}
/* Romberg::integrate, ref: src/utils/NCRomberg.cc:62-146 */
static double gos_romberg(const orc_sc* S, double sasg, double cacg, double acc, double a, double b)
{
  double h = (b - a), f[17];
  gos_many(S, sasg, cacg, f, 17, a, h*0.0625);
  h *= 0.5;
  double R00 = (f[0] + f[16])*h;
  double R10 = h*f[8] + 0.5*R00;
  double R11 = (4./3.)*R10 + (-1./3.)*R00;
  h *= 0.5;
  double R20 = h*(f[4]+f[12]) + 0.5*R10;
  double R21 = (4./3.)*R20 + (-1./3.)*R10;
  double R22 = (16./15.)*R21 + (-1./15.)*R11;
  h *= 0.5;
  double R30 = h*((f[2]+f[6])+(f[10]+f[14])) + 0.5*R20;
  double R31 = (4./3.)*R30 + (-1./3.)*R20;
  double R32 = (16./15.)*R31 + (-1./15.)*R21;
  double R33 = (64./63.)*R32 + (-1./63.)*R22;
  h *= 0.5;
  double R40 = h*(((f[1]+f[3])+(f[5]+f[7]))+((f[9]+f[11])+(f[13]+f[15]))) + 0.5*R30;
  double R41 = (4./3.)*R40 + (-1./3.)*R30;
  double R42 = (16./15.)*R41 + (-1./15.)*R31;
  double R43 = (64./63.)*R42 + (-1./63.)*R32;
  double R44 = (256./255.)*R43 + (-1./255.)*R33;
  if (gos_accept(acc, 4, R33, R44)) return R44;
  double c5 = gos_manysum(S, sasg, cacg, 16, a + h*0.5, h);
  h *= 0.5;
  double R50 = h*c5 + 0.5*R40;
  double R51 = (4./3.)*R50 + (-1./3.)*R40;
  double R52 = (16./15.)*R51 + (-1./15.)*R41;
  double R53 = (64./63.)*R52 + (-1./63.)*R42;
  double R54 = (256./255.)*R53 + (-1./255.)*R43;
  double R55 = (1024./1023.)*R54 + (-1./1023.)*R44;
  if (gos_accept(acc, 5, R44, R55)) return R55;
  double c1[16], c2[16], *rp = c1, *r = c2;
  rp[0] = R50; rp[1] = R51; rp[2] = R52; rp[3] = R53; rp[4] = R54; rp[5] = R55;
  unsigned nj = 16;
  for (unsigned i = 6; i < 16; ++i) {
    double hh = h;
    h *= 0.5; nj *= 2;
    double c = gos_manysum(S, sasg, cacg, nj, a + h, hh);
    r[0] = h*c + 0.5*rp[0];
    double nk = 1.;
    for (unsigned j = 0; j < i; ++j) { nk *= 4.0; r[j+1] = (nk*r[j] - rp[j])/(nk - 1.0); }
    if (gos_accept(acc, i, rp[i-1], r[i])) return r[i];
    double* t = rp; rp = r; r = t;
  }
  return rp[15];
}
/* GaussOnSphere::circleIntegralSlow / circleIntegral, ref: NCGaussOnSphere.cc:380-433, NCGaussOnSphere.hh:194-208 */
static double gos_circle_slow(const orc_sc* S, double cg, double sg, double ca, double sa)
{
  double sasg = sa*sg, cacg = ca*cg, cd = cacg + sasg;
  if (cd <= S->cta) return 0.0;
  if (sasg < 1e-14) return SC_2PI*sa*gos_cosx(S, ca);
  double cos_tmax = (S->cta - cacg)/sasg;
  double tmax = (cos_tmax <= -1.0 ? ORC_PI : acos(orc_min(1.0, cos_tmax)));
  if (tmax <= 1e-12) return 0.0;
  double acc = S->numint_accuracy;
  if (tmax < 10*SC_ARCSEC) { acc = orc_max(acc, 1e-6); if (tmax < SC_ARCSEC) { acc = orc_max(acc, 1e-5); if (tmax < 0.1*SC_ARCSEC) acc = orc_max(acc, 1e-4); } }
  return 2.0*sa*gos_romberg(S, sasg, cacg, acc, 0, tmax);
}
static double gos_circle(const orc_sc* S, double cg, double sg, double ca, double sa)
{
  double sasg = sa*sg, cacg = ca*cg, cd = cacg + sasg;
  if (cd > S->cta && sasg >= 1e-14 && S->k2 > S->k1*sasg + cacg) return lut_eval(&S->sofcosd, cd)*sqrt(sa/sg);
  return gos_circle_slow(S, cg, sg, ca, sa);
}

/* GaussMos_cacheRound / SCBragg_cacheRound, ref: src/phys_utils/NCGaussMos.cc:28-36, src/scbragg/NCSCBragg.cc:224-230 */
static double gm_round(double x) { return floor(orc_max(x, 1e-15)*1e15 + 0.5)*1e-15; }
static double sc_round(double x) { return floor(x*1e15 + 0.5)*1e-15; }

/* GaussMos::InteractionPars, ref: NCGaussMos.hh:185-205, NCGaussMos.cc:252-280 (with the reference's caching) */
typedef struct { double Q, spt, cpt, wl, wl3, inv2dsp, cptsq, Qprime, xsfact; } ipars;
static void ip_set(ipars* ip, double wl_raw, double inv2dsp_raw, double xsfact)
{
  ip->xsfact = xsfact*0.5;
  double wl = gm_round(wl_raw), inv2dsp = gm_round(inv2dsp_raw);
  if (wl == ip->wl) {
    if (inv2dsp == ip->inv2dsp) { ip->Q = (ip->Qprime > 0.0 ? ip->Qprime*ip->xsfact : -1); return; }
  } else { ip->wl = wl; ip->wl3 = wl*wl*wl; }
  ip->inv2dsp = inv2dsp;
  ip->spt = wl*inv2dsp;
  ip->cptsq = 1 - ip->spt*ip->spt;
  ip->Q = ip->Qprime = ip->cpt = -1;
}
/* calcRawCrossSectionValue (+Init), ref: NCGaussMos.hh:248-258, NCGaussMos.cc:116-145 */
static double gm_rawxs(const orc_sc* S, ipars* ip, double cosang)
{
  cosang = orc_clamp(cosang, -1.0, 1.0);
  if (!(ip->Q > 0.)) {
    if (ip->Qprime == -1) {
      ip->cpt = sqrt(ip->cptsq);
      double t = ip->cpt*ip->spt;
      if (t > 0) ip->Qprime = ip->wl3/t;
      else ip->Qprime = (ip->spt > 0.5 && ip->xsfact) ? -2.0 : 0.0;
    }
    if (ip->Qprime > 0.) ip->Q = ip->Qprime*ip->xsfact;
    else return ip->Qprime ? HUGE_VAL : 0.0;
  }
  double sinang = sqrt(1.0 - cosang*cosang);
  return ip->Q*gos_circle(S, cosang, sinang, ip->spt, ip->cpt);
}

/* per-(E,dir) cache of SCBragg, ref: NCSCBragg.cc:68-79 */
typedef struct { double* commul; orc_vec* normal; double* inv2d; int n, cap; double wl; } sc_cache;
static void cache_push(sc_cache* c, double v, orc_vec nrm, double i2d)
{
  if (c->n == c->cap) {
    c->cap = c->cap ? 2*c->cap : 64;
    c->commul = (double*)realloc(c->commul, sizeof(double)*c->cap);
    c->normal = (orc_vec*)realloc(c->normal, sizeof(orc_vec)*c->cap);
    c->inv2d = (double*)realloc(c->inv2d, sizeof(double)*c->cap);
  }
  c->commul[c->n] = v; c->normal[c->n] = nrm; c->inv2d[c->n] = i2d; ++c->n;
}
static void cache_free(sc_cache* c) { free(c->commul); free(c->normal); free(c->inv2d); memset(c, 0, sizeof(*c)); }

/* SCBragg::pimpl::updateCache + GaussMos::calcCrossSections, ref: NCSCBragg.cc:233-275, NCGaussMos.cc:147-194 */
static void sc_update_cache(const orc_sc* S, sc_cache* c, double ekin_raw, orc_vec dir)
{
  double ekin = sc_round(ekin_raw);
  dir = vnorm(dir);
  c->n = 0;
  c->wl = ekin ? sqrt(ORC_WL2EKIN/ekin) : HUGE_VAL;    /* ekin2wl, NCDefs.hh:840-845 */
  if (c->wl == 0) return;
  double cutoff = (1.0 - 2*DBL_EPSILON)/c->wl;
  ipars ip; memset(&ip, 0, sizeof(ip)); ip.wl = -1.0; ip.inv2dsp = -1.0;
  for (int f = 0; f < S->nfam; ++f) {
    if (S->fam_inv2d[f] >= cutoff) break;
    ip_set(&ip, c->wl, S->fam_inv2d[f], S->fam_xsfact[f]);
    double xsoffset = c->n ? c->commul[c->n-1] : 0.0, xssum = 0.0;
    double cptsq = ip.cptsq, cta = S->cta;
    int n0 = (int)S->fam_first[f], n1 = (int)S->fam_first[f+1];
    for (int in = n0; in < n1; ++in) {
      orc_vec nrm = { S->normals[3*in], S->normals[3*in+1], S->normals[3*in+2] };
      double dot = vdot(nrm, dir);
      double sd = (1.0 - dot*dot)*cptsq;
      double ds = dot*ip.spt;
      double A0 = orc_max(0.0, cta - fabs(ds));
      if (sd <= A0*A0) continue;
      double Am = orc_max(0.0, cta - ds);
      if (sd > Am*Am) {
        double xs = gm_rawxs(S, &ip, dot);
        if (xs) { orc_vec m = { -nrm.x, -nrm.y, -nrm.z }; cache_push(c, xsoffset + (xssum += xs), m, ip.inv2dsp); }
      }
      double Ap = orc_max(0.0, cta + ds);
      if (sd > Ap*Ap) {
        double xs = gm_rawxs(S, &ip, -dot);
        if (xs) cache_push(c, xsoffset + (xssum += xs), nrm, ip.inv2dsp);
      }
    }
  }
}

/* SCBragg::crossSection, ref: NCSCBragg.cc:295-302 */
double orc_sc_xs(const orc_sc* S, double ekin, orc_vec dir, int* nentries)
{
  *nentries = 0;
  if (ekin <= S->threshold_ekin) return 0.0;
  sc_cache c; memset(&c, 0, sizeof(c));
  sc_update_cache(S, &c, ekin, dir);
  double xs = c.n ? c.commul[c.n-1] : 0.0;
  *nentries = c.n;
  cache_free(&c);
  return xs;
}

/* randPointOnUnitCircle, ref: src/utils/NCRandUtils.cc:98-112 */
static void rand_circle(orc_rng* rng, double* x, double* y)
{
  double a, b, m2;
  do { a = -1.0 + orc_rand(rng)*2.0; b = -1.0 + orc_rand(rng)*2.0; m2 = a*a + b*b; } while (!orc_in(0.001, 1.0, m2));
  double m = 1.0/sqrt(m2);
  *x = a*m; *y = b*m;
}
/* GaussOnSphere::genPointOnCircle, ref: NCGaussOnSphere.cc:435-508; coinflip of a custom stream = generate()>0.5 (src/interfaces/NCRNG.cc:35-38) */
static int gos_gen_point(const orc_sc* S, orc_rng* rng, double cg, double sg, double ca, double sa, double* ct, double* st)
{
  double sasg = sa*sg, cacg = ca*cg, cd = cacg + sasg;
  if (cd <= S->cta) return 0;
  if (sasg < 1e-14) { if (sa < 1e-7) return 0; rand_circle(rng, ct, st); return 1; }
  double cos_tmax = (S->cta - cacg)/sasg;
  if (cos_tmax >= 1.0) return 0;
  double tmax = (cos_tmax <= -1.0 ? ORC_PI : acos(cos_tmax));
  double dmax = gos_cosx_inrange(S, cd)*1.00000001;
  int tries = 1001;
  while (--tries) {
    *ct = cos_mpipi(orc_rand(rng)*tmax);
    double dens = gos_cosx_inrange(S, sasg*(*ct) + cacg);
    if (dens > dmax*orc_rand(rng)) break;
  }
  if (tries <= 0) return 0;
  *st = sqrt(1.0 - (*ct)*(*ct));
  *st = (orc_rand(rng) > 0.5) ? *st : -*st;
  return 1;
}
/* PhiRot::rotateVectorAroundAxis, ref: include/NCrystal/internal/utils/NCRotMatrix.hh:172-191 */
static orc_vec phirot(double cp, double sp, orc_vec v, orc_vec axis)
{
  orc_vec axv = vcross(axis, v);
  double adv = vdot(axis, v), k = sp*1.0, k2;
  orc_vec r = { v.x*cp, v.y*cp, v.z*cp };
  r.x += axv.x*k; r.y += axv.y*k; r.z += axv.z*k;
  k2 = adv*(1.0 - cp);
  r.x += axis.x*k2; r.y += axis.y*k2; r.z += axis.z*k2;
  return r;
}
/* rotateToFrame, ref: src/utils/NCRotMatrix.cc:84-147 */
static orc_vec rotate_to_frame(double sinab, double cosab, orc_vec a, orc_vec b, orc_vec v, orc_rng* rng)
{
  if (fabs(sinab) < 1e-10) {
    double pc = b.z, ps = -sqrt(1.0 - b.z*b.z), rc, rs;
    orc_vec axis = { b.y, -b.x, 0. };
    double m2 = vmag2(axis);
    if (m2 > 1e-12) { double f = 1.0/sqrt(m2); axis.x *= f; axis.y *= f; axis.z *= f; v = phirot(pc, ps, v, axis); }
    else if (b.z < 0.0) v.z *= -1.0;
    rand_circle(rng, &rc, &rs);
    v = phirot(rc, rs, v, b);
    return vnorm(v);
  }
  double s = 1.0/sinab;
  orc_vec c1 = { b.x*(-cosab), b.y*(-cosab), b.z*(-cosab) };
  c1.x += a.x; c1.y += a.y; c1.z += a.z;
  c1.x *= s; c1.y *= s; c1.z *= s;
  orc_vec c2 = vcross(b, a);
  c2.x *= s; c2.y *= s; c2.z *= s;
  orc_vec r = { v.x*c1.x + v.y*c2.x + v.z*b.x, v.x*c1.y + v.y*c2.y + v.z*b.y, v.x*c1.z + v.y*c2.z + v.z*b.z };
  return vnorm(r);
}
/* GaussMos::genScat, ref: NCGaussMos.cc:196-250 */
static orc_vec gm_genscat(const orc_sc* S, orc_rng* rng, orc_vec pn, double pinv2d, double wl_raw, orc_vec indir)
{
  double wl = gm_round(wl_raw), inv2d = gm_round(pinv2d);
  double sb = wl*inv2d;
  if (sb == 0.) return indir;
  double ca = sb, sa = sqrt(1.0 - ca*ca);
  double cg = orc_clamp(-vdot(indir, pn), -1.0, 1.0), sg = sqrt(1.0 - cg*cg), ct, st;
  if (!gos_gen_point(S, rng, cg, sg, ca, sa, &ct, &st)) return indir;
  double s2a = 2*sa*ca, c2a = ca*ca - sa*sa;
  orc_vec out = { s2a*ct, s2a*st, c2a }, mind = { -indir.x, -indir.y, -indir.z };
  out = rotate_to_frame(sg, cg, pn, mind, out, rng);
  return vnorm(out);
}
/* SCBragg::sampleScatter + pimpl::genScat, ref: NCSCBragg.cc:277-288,304-323 */
void orc_sc_sample(const orc_sc* S, double ekin, orc_vec indir, int nentries, double total, orc_rng* rng, orc_vec* out)
{
  (void)nentries; (void)total;
  *out = indir;
  if (ekin <= S->threshold_ekin) return;
  sc_cache c; memset(&c, 0, sizeof(c));
  sc_update_cache(S, &c, ekin, indir);
  if (c.n == 0 || c.commul[c.n-1] <= 0.0) { cache_free(&c); return; }
  int idx = (c.n == 1 ? 0 : orc_pick(orc_rand(rng), c.commul, c.n));
  *out = gm_genscat(S, rng, c.normal[idx], c.inv2d[idx], c.wl, vnorm(indir));
  cache_free(&c);
}

/* randDirectionGivenScatterMu, ref: src/utils/NCRandUtils.cc:51-96 */
orc_vec orc_rand_dir_given_mu(orc_rng* rng, double mu, orc_vec indir)
{
  double m2 = vmag2(indir);
  double invm = (fabs(m2 - 1.0) < 1e-14 ? 1.0 : 1.0/sqrt(m2));
  orc_vec u = { indir.x*invm, indir.y*invm, indir.z*invm }, tmp = { 0, 0, 0 };
  double tm2 = 0.0;
  do {
    double x0 = 2.0*orc_rand(rng) - 1.0, x1 = 2.0*orc_rand(rng) - 1.0, s = x0*x0 + x1*x1;
    if (s < 1.0) {
      double t = 2.0*sqrt(1.0 - s);
      orc_vec d = { x0*t, x1*t, 1.0 - 2.0*s };
      tmp = vcross(d, u);
      tm2 = vmag2(tmp);
    }
  } while (tm2 < 0.001);
  u.x *= mu; u.y *= mu; u.z *= mu;
  double f = sqrt((1 - mu*mu)/tm2);
  u.x += tmp.x*f; u.y += tmp.y*f; u.z += tmp.z*f;
  return u;
}

/* ProcComposition::crossSection (updateCacheAnisotropic), ref: src/interfaces/NCProcImpl.cc:206-249,340-351 */
double orc_xs(const orc_material* M, double ekin, orc_vec dir, double* cumul, int* aux, double* sc_total)
{
  if (!orc_domain_contains(M->dom_lo, M->dom_hi, ekin)) return 0.0;
  double tot = 0.0;
  for (int i = 0; i < M->ncomp; ++i) {
    const orc_comp* c = &M->comp[i];
    int a = -1;
    double xs = 0.0;
    if (orc_domain_contains(c->dom_lo, c->dom_hi, ekin)) {
      if (c->kind == NCB_KIND_SCBRAGG) { xs = orc_sc_xs(&M->sc, ekin, dir, &a); if (sc_total) *sc_total = xs; }
      else xs = orc_comp_xs_iso(M, i, ekin, &a);
    }
    tot += c->scale*xs;
    if (cumul) cumul[i] = tot;
    if (aux) aux[i] = a;
  }
  return tot;
}
/* ProcComposition::sampleScatter, ref: NCProcImpl.cc:364-377; isotropic leaves: ScatterIsotropicMat::sampleScatter :29-37 */
void orc_sample(const orc_material* M, double ekin, orc_vec dir, orc_rng* rng, double* eout, orc_vec* out, int* err)
{
  *eout = ekin; *out = dir;
  if (!orc_domain_contains(M->dom_lo, M->dom_hi, ekin)) return;
  double cumul[ORC_MAXCOMP], sc_total = 0.0; int aux[ORC_MAXCOMP];
  orc_xs(M, ekin, dir, cumul, aux, &sc_total);
  int ich = (M->ncomp == 1 ? 0 : orc_pick(orc_rand(rng), cumul, M->ncomp));
  if (M->comp[ich].kind == NCB_KIND_SCBRAGG) { orc_sc_sample(&M->sc, ekin, dir, aux[ich], sc_total, rng, out); return; }
  double mu;
  orc_comp_sample_iso(M, ich, aux[ich], ekin, rng, eout, &mu, err);
  if (*err & (ORC_ERR_KIN | ORC_ERR_OUTER | ORC_ERR_INNER | ORC_ERR_DISCARD)) { out->x = out->y = out->z = 0.0; *eout = -1.0; return; }
  *out = orc_rand_dir_given_mu(rng, mu, dir);
}
