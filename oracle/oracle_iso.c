/* oracle/oracle_iso.c -- TEST INFRASTRUCTURE (see oracle_common.h).  Isotropic leaves:
 * PowderBragg, ElIncScatter, FreeGas, SABScatter (incl. the sequential SABIntegrator table
 * builder) and the ProcComposition weighted sum / component choice. */
#include "oracle_common.h"

/* ------------------------------------------------------------------ PowderBragg */
/* findLastValidPlaneIdx, ref: src/powderbragg/NCPowderBragg.cc:152-163 */
static int pb_last_valid(const orc_pb* T, double ekin) { return orc_upper_bound(T->e2d, 1, T->n, ekin) - 1; }

/* crossSectionIsotropic, ref: NCPowderBragg.cc:166-176 (inv_ekin cached at :42-50) */
static double pb_xs(const orc_pb* T, double ekin, int* idx)
{
  *idx = -1;
  if (ekin < T->threshold || !isfinite(ekin)) return 0.0;
  *idx = pb_last_valid(T, ekin);
  double inv_ekin = 1.0 / ekin;
  return T->fdm[*idx] * inv_ekin;
}

/* genScatterMu / sampleScatterIsotropic, ref: NCPowderBragg.cc:178-216 */
static void pb_sample(const orc_pb* T, double ekin, int idx, orc_rng* rng, double* eout, double* mu)
{
  *eout = ekin;
  if (ekin < T->threshold || !isfinite(ekin)) { *mu = 1.0; return; }
  if (idx < 0) idx = pb_last_valid(T, ekin);
  int j = orc_lower_bound(T->fdm, 0, idx, orc_rand(rng) * T->fdm[idx]);
  double sin_theta_bragg_squared = T->e2d[j] / ekin;
  *mu = 1.0 - 2.0 * sin_theta_bragg_squared;
}

/* ------------------------------------------------------------------ ElIncXS */
/* eval_1mexpmtdivt, ref: src/phys_utils/NCElIncXS.cc:34-51 */
static double elinc_f(double t)
{
  if (t < 0.01) return (1 + t * (-0.5 + t * 0.16666666666666666666666666666666666666666667 * (1. - 0.25 * t)));
  if (t > 24.0) return 1.0 / t;
  t = -t;
  return expm1(t) / t;
}
/* evaluate / evalXSContribsCommul, ref: NCElIncXS.cc:117-141 */
static double elinc_xs(const orc_elinc* T, double ekin, double* contribs)
{
  const double kkk = 16.0 * ORC_PISQ * ORC_EKIN2WLSQINV;
  double e = kkk * ekin, xs = 0.0;
  for (int i = 0; i < T->n; ++i) { xs += T->bixs[i] * elinc_f(T->msd[i] * e); if (contribs) contribs[i] = xs; }
  return xs;
}
/* exp_smallarg_approx, ref: include/NCrystal/internal/utils/NCMath.hh:434-440 */
static double exp_small(double x)
{
  return 1.0+x*(1+x*(0.5+x*(0.16666666666666666666666666666666666667+x*(0.04166666666666666666666666666666666667
         +x*(0.00833333333333333333333333333333333333+x*(0.00138888888888888888888888888888888889
         +x*0.00019841269841269841269841269841269841))))));
}
/* sampleMuMonoAtomic, ref: NCElIncXS.cc:81-115 */
static double elinc_mu_mono(orc_rng* rng, double ekin, double msd)
{
  const double kkk = 8.0 * ORC_PISQ * ORC_EKIN2WLSQINV;
  double twoksq = kkk * ekin;
  double a = twoksq * msd;
  if (a < 0.01) {
    double maxval = exp_small(a);
    for (;;) {
      double mu = orc_rand(rng) * 2.0 - 1.0;
      if (orc_rand(rng) * maxval < exp_small(a * mu)) return mu;
    }
  }
  return orc_clamp(log1p(orc_rand(rng) * expm1(2.0 * a)) / a - 1.0, -1.0, 1.0);
}
/* EPointAnalysis::sampleMu, ref: NCElIncXS.cc:178-190 */
static double elinc_sample_mu(const orc_elinc* T, double ekin, orc_rng* rng)
{
  if (T->n == 1) return elinc_mu_mono(rng, ekin, T->msd[0]);
  double contribs[ORC_MAXEL];
  elinc_xs(T, ekin, contribs);
  int c = orc_pick(orc_rand(rng), contribs, T->n);
  return elinc_mu_mono(rng, ekin, T->msd[c]);
}

/* ------------------------------------------------------------------ kinematics */
/* getAlphaLimits, ref: include/NCrystal/internal/phys_utils/NCKinUtils.hh:85-124 */
static void alpha_limits(double ediv, double beta, double* amin, double* aplus)
{
  double kk = ediv + beta;
  if (!(kk >= 0.0)) { *amin = 1.0; *aplus = -1.0; return; }
  double a = kk + ediv;
  double b = 2.0 * sqrt(ediv * kk);
  if (fabs(beta) < 0.01 * ediv) {
    double x = beta / ediv;
    const double c9 = -715./32768., c8 = 429./16384., c7 = -33./1024., c6 = 21./512., c5 = -7./128., c4 = 5./64., c3 = -1./8., c2 = 1./4.;
    *amin = beta*x*(c2+x*(c3+x*(c4+x*(c5+x*(c6+x*(c7+x*(c8+x*c9)))))));
  } else {
    *amin = orc_max(0.0, a - b);
  }
  *aplus = a + b;
}
/* muIsotropicAtBeta, ref: NCKinUtils.hh:64-70 */
static int mu_iso_at_beta(double beta, double ediv) { const double lim = -1.0 + 1e-14; return beta <= ediv * lim; }
/* convertAlphaBetaToDeltaEMu, ref: src/phys_utils/NCKinUtils.cc:25-56 */
static void ab_to_demu(double alpha, double beta, double ekin, double kT, double* dE, double* mu, int* err)
{
  *dE = beta * kT;
  double ef = ekin + *dE;
  double denom = 2.0 * sqrt(ekin * ef);
  if (!denom) { *err |= ORC_ERR_KIN; *mu = -999.0; return; }
  orc_ssum s = {0, 0};
  orc_ssum_add(&s, ekin); orc_ssum_add(&s, ef); orc_ssum_add(&s, -alpha * kT);
  *mu = orc_clamp(orc_ssum_get(&s) / denom, -1.0, 1.0);
}

/* ------------------------------------------------------------------ free gas */
/* FreeGasXSProvider::evalXSShapeASq, ref: src/phys_utils/NCFreeGasUtils.cc:64-83 */
static double fg_shape(double a2)
{
  if (a2 > 36.0) return 1.0 + 0.5 / a2;
  double a = sqrt(a2);
  if (a < 0.1) {
    if (a == 0.0) return HUGE_VAL;
    const double c1 = 2.0/3.0, c2 = 1.0/15.0, c3 = 1.0/105.0, c4 = 1.0/756.0, c5 = 1.0/5940.0;
    return ORC_INVSQRTPI * (2.0 / a + a * (c1 - a2*(c2 - a2*(c3 - a2*(c4 - a2*c5)))));
  }
  double inva = 1.0 / a;
  return (1.0 + 0.5*inva*inva) * erf(a) + ORC_INVSQRTPI * exp(-a2) * inva;
}
static double fg_xs(const orc_fg* T, double ekin) { return T->sigma_free * fg_shape(T->ca * ekin); }

/* erfc lookup table, ref: NCFreeGasUtils.cc:92-150 */
#define ORC_LUTLEN 1103
static double g_erfc_lut[ORC_LUTLEN];
static int g_erfc_lut_ok = 0;
static void erfc_lut_init(void)
{
  if (g_erfc_lut_ok) return;
  const double lo = -2.0, hi = 9.0;
  const int nedges = ORC_LUTLEN - 2;
  double interval = (hi - lo) / (nedges - 1);   /* linspace, src/utils/NCMath.cc:68-81 */
  g_erfc_lut[0] = 2.0;
  for (int i = 0; i < nedges - 1; ++i) g_erfc_lut[1 + i] = erfc(lo + i * interval);
  g_erfc_lut[nedges] = erfc(hi);
  g_erfc_lut[ORC_LUTLEN - 1] = 0.0;
  g_erfc_lut_ok = 1;
}
static void erfc_bounds(double x, double* lb, double* ub)
{
  const double lo = -2.0, hi = 9.0;
  const double binw = (hi - lo) / (ORC_LUTLEN - 3), invbinw = (ORC_LUTLEN - 3) / (hi - lo);
  double xs = orc_clamp(x, lo - 0.5*binw, hi + 0.5*binw);
  int b = (int)(1.0 + (xs - lo) * invbinw);
  if (b > ORC_LUTLEN - 1) b = ORC_LUTLEN - 1;
  if (b < 0) b = 0;
  *lb = g_erfc_lut[b + 1] * 0.99999999;
  *ub = g_erfc_lut[b] * 1.00000001;
}
/* erfcdiff, ref: src/utils/NCMath.cc:337-390 */
static double erfcdiff_nt(double a, double b)
{
  if (b < 0) { double t; b = -b; a = -a; t = a; a = b; b = t; }
  double ea = a > 27.3 ? 0.0 : erfc(a);
  if (b > a + 4.0 && (a >= 4 || (a < 0.0 && b > 6.0))) return ea;
  double eb = b > 27.3 ? 0.0 : erfc(b);
  return ea - eb;
}
static double erfcdiff(double a, double b)
{
  if (orc_max(fabs(a), fabs(b)) < 0.32) {
    const double c1 = -2.0*ORC_INVSQRTPI, c3 = 2.0*ORC_INVSQRTPI/3.0, c5 = -0.2*ORC_INVSQRTPI, c7 = ORC_INVSQRTPI/21.0,
                 c9 = -ORC_INVSQRTPI/108.0, c11 = ORC_INVSQRTPI/660.0, c13 = -ORC_INVSQRTPI/4680.0, c15 = ORC_INVSQRTPI/37800.0;
    double a2 = a*a, b2 = b*b;
    double ta = a * a2 * (c3 + a2 * (c5 + a2 * (c7 + (a2 * (c9 + a2 * (c11 + a2 * (c13 + a2 * c15)))))));
    double tb = b * b2 * (c3 + b2 * (c5 + b2 * (c7 + (b2 * (c9 + b2 * (c11 + b2 * (c13 + b2 * c15)))))));
    return c1*(a-b) + (ta - tb);
  }
  return a > b ? -erfcdiff_nt(b, a) : erfcdiff_nt(a, b);
}
/* erfc_rescaled, ref: NCMath.cc:393-415 */
static double erfc_rescaled(double x, double b)
{
  if (b < -745.1) return 0.0;
  if ((x < 23.0 && fabs(b) < 700) || x < 5) return exp(b) * erfc(x);
  double bxx = b - x*x;
  if (bxx < -745.1) return 0.0;
  const double c3 = -0.5, c5 = 0.75, c7 = -1.875, c9 = 6.5625, c11 = -29.53125;
  double y = 1/x, y2 = y*y;
  return ORC_INVSQRTPI*exp(bxx)*(y+y2*(c3+y2*(c5+y2*(c7+y2*(c9+y2*c11)))));
}
/* RandExpIntervalSampler, ref: include/NCrystal/internal/utils/NCRandUtils.hh:180-215 */
typedef struct { double a, c1, c2; } expsampler;
static void es_set(expsampler* s, double a, double b, double c) { s->a = a; s->c1 = -1.0/c; s->c2 = expm1(-c*(b-a)); }
static double es_sample(const expsampler* s, orc_rng* rng) { return s->a + s->c1 * log(1.0 + orc_rand(rng) * s->c2); }

/* randExpDivSqrt, ref: src/utils/NCRandUtils.cc:224-380 */
static double rand_exp_div_sqrt(orc_rng* rng, double c, double a, double b)
{
  double A = c*a;
  if (A > 0.1) {
    double U = c*(b-a), invA = 1.0/A;
    expsampler es; es_set(&es, 0, U, 1.0);
    for (;;) {
      double ugen = es_sample(&es, rng);
      double R = orc_rand(rng);
      if ((1.0 + ugen*invA)*R*R < 1.0) return orc_clamp((ugen + A)/c, a, b);
    }
  } else {
    const double Ulim = 16.1180956509583;
    double U = orc_min(c*(b-a), Ulim);
    double B = U + A;
    if (!(B > A)) return a;
    double sqrtA = sqrt(A), sqrtB = sqrt(B);
    double dsq = sqrtB - sqrtA, twosqrtA = 2*sqrtA, ugen;
    for (;;) {
      double T = orc_rand(rng)*dsq;
      ugen = T*(T + twosqrtA);
      double Racc = orc_rand(rng);
      if (ugen < 2.0) {
        const double c1 = -1.0, c2 = 1.0/2.0, c3 = -1.0/6.0, c4 = 1.0/24.0, c5 = -1.0/120.0, c6 = 1.0/720.0;
        double t6 = 1.0+ugen*(c1+ugen*(c2+ugen*(c3+ugen*(c4+ugen*(c5+ugen*c6)))));
        if (Racc > t6) continue;
        if (Racc + 0.020221 < t6) break;
      } else {
        if (Racc > 0.135335283236614) continue;
        if (ugen > 4.0 && Racc > 0.0183156388887343) continue;
      }
      if (Racc < exp(-ugen)) break;
    }
    return orc_clamp((ugen + A)/c, a, b);
  }
}

/* f_eval of randExpMInvXMCXDivSqrtX, ref: NCFreeGasUtils.cc:314-322 */
static double fg_feval(double xmax, double c, double x)
{
  double ea = (x-xmax)/(x*xmax) - c*(x-xmax);
  if (ea >= 706.0) return 1.0;
  return ea < -745.1 ? 0.0 : exp(ea)*sqrt(xmax/x);
}
/* randExpMInvXMCXDivSqrtX, ref: NCFreeGasUtils.cc:237-490 */
static double rand_expminvx(orc_rng* rng, double c, double xm, double xp)
{
  if (xp == xm) return xm;
  double sqrtc = sqrt(c), invsqrtc = 1/sqrtc;
  double xpeak = (c > 1e-5 ? (c > 1e200 ? invsqrtc : (sqrt(16.0*c+1.0)-1.0)/(4.0*c))
                           : (2.0-c*(8.0-c*(64.0-c*(640.0-c*7168.0)))));
  if (xpeak == 0.0) return xm > 0.0 ? xm : orc_min(DBL_MIN, xp);
  double xmax = (xm > xpeak ? xm : orc_min(xp, xpeak));
  if (!(xmax > 0.0)) return xm;
  double xlarge = orc_max(5.0/sqrt(c), 2*xpeak);
  double xsmall = orc_min(0.2/sqrt(c), 0.5*xpeak);
  if (xp > xlarge) xp = orc_min(xp, orc_max(xm, xlarge) + 15.0/c);
  if (xm < xsmall) { double xsm = orc_min(xp, xsmall); xm = orc_max(xm, xsm/(1 + 30.0*xsm)); }
  if ((xm = orc_max(DBL_MIN, orc_max(DBL_MIN/xp, xm))) >= xp) return xp;
  const double fcut = 1e-9;
  if (xp < xpeak) {
    for (;;) { double xn = xp - 0.01*(xp-xm); if (fg_feval(xmax, c, xn) >= fcut) break; xm = xn; }
  }
  double pflat = -1.0, xswitch = -1.0, area_right = -1.0;
  if (xm >= xlarge) { pflat = 0.0; xswitch = xm; }
  else if (c > 25 || xp <= xlarge) { pflat = 1.0; xswitch = xp; }
  else {
    xswitch = xlarge;
    double area_left = (xswitch - xm);
    double B = c*xmax + 1/xmax - 1/xp;
    area_right = (erfc_rescaled(sqrtc*sqrt(xswitch), B) - erfc_rescaled(sqrtc*sqrt(xp), B))*sqrt(ORC_PI*(xmax/c));
    pflat = area_left/(area_left + area_right);
  }
  int always_left = (pflat > 1.0 - fcut), always_right = (pflat < fcut);
  int single_side = (always_left || always_right);
  if (!single_side && fg_feval(xmax, c, xswitch) < fcut*1.1) { pflat = 1.0; area_right = 0.0; xp = xswitch; always_left = 1; single_side = 0; }
  for (;;) {
    int do_flat = single_side ? always_left : (orc_rand(rng) < pflat);
    if (do_flat) {
      double dx = xswitch - xm;
      double xgen = xm + orc_rand(rng)*dx;
      double Racc = orc_rand(rng);
      /* the reference's xthr_low/xthr_up are loop-local (:434-441): equal to (xm,xswitch) here */
      if (!orc_in(xm, xswitch, xgen) && Racc > 0.05) continue;
      double fval = fg_feval(xmax, c, xgen);
      if (fval < 0.05) {
        if (fval < fcut) {
          if (xgen < xmax) xm = xgen; else xswitch = xgen;
          dx = xswitch - xm;
          if (!single_side) {
            pflat = dx/(dx + area_right);
            always_left = (pflat > 1.0 - fcut); always_right = (pflat < fcut);
            single_side = (always_left || always_right);
          }
          continue;
        }
      }
      if (Racc <= fval) return xgen;
    } else {
      double xgen = rand_exp_div_sqrt(rng, c, xswitch, xp);
      if (orc_rand(rng) < exp((xgen-xp)/(xgen*xp))) return xgen;
    }
  }
}

/* FGEvalBetaDistHelper, ref: NCFreeGasUtils.cc:153-233 */
typedef struct { double beta, normfact, expmbeta, k11, k12, k21, k22; } fgdist;
static void fgdist_init(fgdist* d, double c, double invA, double sqrtAc, double beta, double normfact)
{
  d->beta = beta; d->normfact = normfact; d->expmbeta = -1.0;
  double eps = beta/c;
  double s1pe = sqrt(1+eps);
  double S = (beta < 0.0 ? -1.0 : 1.0);
  double sep = (eps >= 0.0 ? 1.0 : s1pe);
  double sgp = sqrt(2.0+eps+2.0*s1pe);
  double SP = 0.5*(S+invA), SM = 0.5*(S-invA);
  double ia = invA*sep, ms = -S*sep;
  double SPs = sgp*SP, SMs = sgp*SM;
  d->k11 = sqrtAc*(-ia + SPs);
  d->k12 = sqrtAc*(ms + SPs);
  d->k21 = sqrtAc*(ms + SMs);
  d->k22 = sqrtAc*(ia + SMs);
}
static void fgdist_expmb(fgdist* d) { if (d->expmbeta < 0) d->expmbeta = d->beta < -700.0 ? 0.0 : exp(-d->beta); }
static double fgdist_exact(fgdist* d)
{
  double t1 = erfcdiff(d->k11, d->k12);
  fgdist_expmb(d);
  if (!d->expmbeta) return d->normfact*t1;
  double t2 = erfcdiff(d->k21, d->k22);
  return d->normfact*(t1 + t2*d->expmbeta);
}
static void fgdist_bounds(fgdist* d, double* lb, double* ub)
{
  double l11,u11,l12,u12,l21,u21,l22,u22;
  erfc_bounds(d->k11,&l11,&u11); erfc_bounds(d->k12,&l12,&u12);
  double t1l = l11 - u12, t1u = u11 - l12;
  erfc_bounds(d->k21,&l21,&u21); erfc_bounds(d->k22,&l22,&u22);
  double t2l = l21 - u22, t2u = u21 - l22;
  if (t2u > 0.0) { fgdist_expmb(d); *lb = d->normfact*(t1l + t2l*d->expmbeta); *ub = d->normfact*(t1u + t2u*d->expmbeta); return; }
  *lb = d->normfact*t1l; *ub = d->normfact*t1u;
}

/* FreeGasSampler, ref: NCFreeGasUtils.cc:492-515 */
typedef struct { double c, kT, sqrtAc, invA, Adiv4, normfact, c_real; } fgsampler;
static void fgs_init(fgsampler* s, double ekin, double kT, double mass_amu)
{
  erfc_lut_init();
  s->c = orc_min(1e14, orc_max(1e-10, ekin/kT));
  s->kT = kT;
  s->sqrtAc = sqrt(mass_amu*s->c/ORC_NEUTRON_MASS_AMU);
  double A = (1.0/ORC_NEUTRON_MASS_AMU) * mass_amu;
  s->invA = 1.0/A;
  s->Adiv4 = 0.25*A;
  s->normfact = 0.5/erf(sqrt(s->c*s->invA));
  s->c_real = ekin/kT;
}
typedef struct { double a, b, pdown, pnotclose; expsampler es; int es_valid; } fgoverlay;
static void ov_set(fgoverlay* o, double aaa, double bbb)
{
  const double Tlim = 2.0, k1 = 0.135335283236612691893999494972484403407, k2 = 274./315.;
  o->a = aaa; o->b = bbb;
  double ad = -aaa, af, ac;
  if (bbb <= Tlim) {
    af = 0.0;
    const double c2 = -1./2., c3 = 1./6., c4 = -1./24., c5 = 1./120., c6 = -1./720., c7 = 1./5040.;
    double b = bbb;
    ac = b*(1.0+b*(c2+b*(c3+b*(c4+b*(c5+b*(c6+b*c7))))));
  } else { af = k1 - exp(-bbb); ac = k2; }
  double inv = 1.0/(ad + ac + af);
  o->pdown = ad*inv;
  o->pnotclose = (af + ad)*inv;
  o->es_valid = 0;
}
/* sampleBeta, ref: NCFreeGasUtils.cc:530-849 */
static double fgs_sample_beta(const fgsampler* s, orc_rng* rng)
{
  if (s->c_real > 1e4) {
    double A = 1.0/s->invA, A2 = A*A;
    double thr = 1e4*orc_min(1000.0*A, A2*A2*A2);
    if (s->c_real > thr) {
      double r = (1.0-s->invA)/(1.0+s->invA);
      double elossmax = s->c_real*(1.0 - r*r);
      return -elossmax*orc_rand(rng);
    }
  }
  const double Tlim = 2.0, fcut = 1e-6;
  double aa = orc_max(-s->c_real, -s->c), bb = 13.815510557964274;
  fgdist d;
  if (s->invA <= 1.0/10.0) {
    if (s->c > 10.1) {
      for (;;) {
        double an = aa*0.2;
        if (an > -1e-99) break;
        fgdist_init(&d, s->c, s->invA, s->sqrtAc, an, s->normfact);
        if (fgdist_exact(&d) > fcut) break;
        aa = an;
      }
    }
    for (;;) {
      double bn = bb*0.25, lb, ub;
      if (bn < 1e-99) break;
      fgdist_init(&d, s->c, s->invA, s->sqrtAc, bn, s->normfact);
      fgdist_bounds(&d, &lb, &ub);
      if (ub > fcut) break;
      bb = bn;
    }
  }
  if (!(bb > aa)) return aa;
  fgoverlay ov; ov_set(&ov, aa, bb);
  const double fthr = 0.1;
  double afthr = ov.a, bfthr = ov.b;
  for (;;) {
    double beta, fover;
    double Rsel = orc_rand(rng);
    if (Rsel < ov.pdown) { beta = orc_rand(rng)*ov.a; fover = 1.0; }
    else if (Rsel < ov.pnotclose) {
      if (!ov.es_valid) { es_set(&ov.es, Tlim, ov.b, 1.0); ov.es_valid = 1; }
      beta = es_sample(&ov.es, rng);
      fover = exp(-beta);
    } else {
      double bmax = orc_min(ov.b, Tlim);
      for (;;) {
        beta = orc_rand(rng)*bmax;
        double R0 = orc_rand(rng);
        const double kcheap = 19./45.;
        if (R0 > 1.0 - kcheap*beta) continue;
        const double c1 = -1., c2 = 1./2., c3 = -1./6., c4 = 1./24., c5 = -1./120., c6 = 1./720.;
        fover = 1.0+beta*(c1+beta*(c2+beta*(c3+beta*(c4+beta*(c5+beta*c6)))));
        if (R0 < fover) break;
      }
    }
    double facc = orc_rand(rng)*fover;
    if (facc > fthr && !orc_in(afthr, bfthr, beta)) continue;
    fgdist_init(&d, s->c, s->invA, s->sqrtAc, beta, s->normfact);
    int need_exact = 1;
    double fval = 0.0;
    if (beta > 0) {
      double lb, ub;
      fgdist_bounds(&d, &lb, &ub);
      if (facc <= lb) return beta;
      fval = ub;
      if (facc > ub) need_exact = 0;
    }
    if (need_exact) { fval = fgdist_exact(&d); if (facc < fval) return beta; }
    if (fval < fcut) { if (beta < 0) ov_set(&ov, beta, ov.b); else ov_set(&ov, ov.a, beta); continue; }
    if (fval < fthr) { if (beta < 0) afthr = orc_max(afthr, beta); else bfthr = orc_min(bfthr, beta); }
  }
}
/* sampleAlpha, ref: NCFreeGasUtils.cc:851-935 */
static double fgs_sample_alpha(const fgsampler* s, double beta, orc_rng* rng)
{
  double am, ap;
  if (s->c_real < s->c || mu_iso_at_beta(beta, s->c)) {
    alpha_limits(s->c_real, beta, &am, &ap);
    double alpha = am + orc_rand(rng)*(ap - am);
    return orc_clamp(alpha, am, ap);
  }
  beta = orc_max(-s->c, beta);
  alpha_limits(s->c, beta, &am, &ap);
  if (am == ap) return am;
  double betasq = beta*beta, t = betasq*s->Adiv4, c = 0.0625*betasq;
  if (orc_min(t, c) < 1e-5) {
    double fourA = s->Adiv4*16.0, inv4A = 1.0/fourA;
    double xxm = am*inv4A, xxp = ap*inv4A;
    for (;;) {
      double xx = rand_exp_div_sqrt(rng, 1.0, xxm, xxp);
      double alpha = xx*fourA;
      if (alpha < am || alpha > ap) continue;
      if (alpha*ap*(-log(orc_rand(rng))) >= t*(ap - alpha)) return alpha;   /* randExp, NCRandUtils.hh:175 */
    }
  }
  double invt = 1.0/t;
  double x = rand_expminvx(rng, c, am*invt, ap*invt);
  return orc_clamp(x*t, am, ap);
}
/* sampleAlphaBeta / sampleDeltaEMu, ref: include/NCrystal/internal/phys_utils/NCFreeGasUtils.hh:146-174 */
static void fgs_alpha_beta(const fgsampler* s, orc_rng* rng, double* alpha, double* beta)
{
  *beta = fgs_sample_beta(s, rng);
  if (*beta < -s->c || mu_iso_at_beta(*beta, s->c)) {
    double am, ap;
    alpha_limits(s->c_real, *beta, &am, &ap);
    *alpha = orc_clamp(am + orc_rand(rng)*(ap - am), am, ap);
    return;
  }
  *alpha = fgs_sample_alpha(s, *beta, rng);
}
static void fgs_demu(const fgsampler* s, orc_rng* rng, double* dE, double* mu, int* err)
{
  double beta = fgs_sample_beta(s, rng);
  if (beta <= -s->c || mu_iso_at_beta(beta, s->c)) { *dE = beta*s->kT; *mu = orc_rand(rng)*2.0 - 1.0; return; }
  double alpha = fgs_sample_alpha(s, beta, rng);
  ab_to_demu(alpha, beta, s->c*s->kT, s->kT, dE, mu, err);
}

/* ------------------------------------------------------------------ SAB: cross section */
/* SABXSProvider::crossSection x SABScatter::m_scale, ref: src/sab/NCSABXSProvider.cc:54-95, src/sabscatter/NCSABScatter.cc:87 */
static double sab_xs(const orc_sab* T, double ekin)
{
  int n = T->negrid, iu = orc_upper_bound(T->egrid, 0, n, ekin);
  double xs;
  if (iu == n) xs = T->k_extension/ekin + fg_xs(&T->ext, ekin);
  else if (iu == 0) xs = ekin > 0.0 ? sqrt(T->egrid[0]/ekin)*T->xs[0] : HUGE_VAL;
  else {
    double dXS = T->xs[iu] - T->xs[iu-1], dE = T->egrid[iu] - T->egrid[iu-1];
    xs = T->xs[iu-1] + dXS*(ekin - T->egrid[iu-1])/dE;
  }
  return xs*T->scale;
}

/* ------------------------------------------------------------------ SAB: table builder (sequential) */
/* integrateAlphaInterval_fast, ref: include/NCrystal/internal/sab/NCSABUtils.hh:236-258 */
static double integ_fast(double a1, double s1, double a2, double s2, double l1, double l2)
{
  double da = a2-a1, ps = s1+s2, ds = s2-s1;
  if (orc_min(s1, s2) < 1e-300) return 0.5*da*ps;
  if (fabs(ds) > 0.006*ps) return da*ds/(l2-l1);
  double y = ds/ps, ysq = y*y;
  const double c1 = 0.166666666666666666666666666666666666666666667, c2 = 0.0444444444444444444444444444444444444444444444,
               c3 = 0.0232804232804232804232804232804232804232804233;
  return da*ps*(0.5-ysq*(c1+ysq*(c2+ysq*c3)));
}
/* interpolate_loglin_fallbacklinlin_fast, ref: NCSABUtils.hh:181-200 */
static double interp_fast(double a, double fa, double b, double fb, double x, double la, double lb)
{
  double bma = b-a, mid = 0.5*(b+a);
  int lin = (fa*fb == 0.0);
  if (x < mid) { double r = (x-a)/bma; return lin ? (fa + (fb-fa)*r) : exp(la + (lb-la)*r); }
  double s = (b-x)/bma;
  return lin ? (fb + (fa-fb)*s) : exp(lb + (la-lb)*s);
}
typedef struct { double xs_front, xs_middle, xs_back; unsigned imid_lo, imid_up; double fa, fs, fl, ba, bs, bl; int narrow; } tailed;
static void set_tail(const double* ag, const double* sab, const double* ls, unsigned idx, double alpha, double* ta, double* ts, double* tl)
{
  *ta = alpha;
  *ts = interp_fast(ag[idx], sab[idx], ag[idx+1], sab[idx+1], alpha, ls[idx], ls[idx+1]);
  *tl = log(orc_max(*ts, DBL_MIN));
}
/* createTailedBreakdown, ref: src/sab/NCSABUtils.cc:538-633 */
static tailed tailed_breakdown(const double* ag, int na, const double* sab, const double* ls, const double* cum,
                               double alow, double aupp, unsigned il, unsigned iu)
{
  tailed tb; memset(&tb, 0, sizeof(tb));
  alow = orc_clamp(alow, ag[0], ag[na-1]);
  aupp = orc_clamp(aupp, ag[0], ag[na-1]);
  if (il == iu || alow == aupp) return tb;
  if (il + 1 == iu) {
    tb.narrow = 1;
    set_tail(ag, sab, ls, il, alow, &tb.fa, &tb.fs, &tb.fl);
    set_tail(ag, sab, ls, il, aupp, &tb.ba, &tb.bs, &tb.bl);
    tb.xs_front = integ_fast(tb.fa, tb.fs, tb.ba, tb.bs, tb.fl, tb.bl);
    return tb;
  }
  tb.imid_lo = il; tb.imid_up = iu;
  if (alow >= ag[il]) {
    set_tail(ag, sab, ls, il, alow, &tb.fa, &tb.fs, &tb.fl);
    tb.xs_front = integ_fast(tb.fa, tb.fs, ag[il+1], sab[il+1], tb.fl, ls[il+1]);
    ++tb.imid_lo;
  }
  if (aupp <= ag[iu]) {
    set_tail(ag, sab, ls, iu-1, aupp, &tb.ba, &tb.bs, &tb.bl);
    tb.xs_back = integ_fast(ag[iu-1], sab[iu-1], tb.ba, tb.bs, ls[iu-1], tb.bl);
    --tb.imid_up;
  }
  tb.xs_middle = (tb.imid_up > tb.imid_lo ? cum[tb.imid_up] - cum[tb.imid_lo] : 0.0);
  return tb;
}

/* SABIntegrator::Impl::analyseEnergyPoint (doSampler=true), ref: src/sab/NCSABIntegrator.cc:352-559,
 * with activeGridRanges (src/sab/NCSABUtils.cc:460-536) run in its original sequential form
 * (iterators hinted from the previous beta row). */
static int analyse_epoint(const orc_sab* T, double ekin, orc_epoint* ep)
{
  const int na = T->nalpha, nb = T->nbeta;
  const double* ag = T->alpha; const double* bg = T->beta;
  memset(ep, 0, sizeof(*ep));
  ep->first_bin = 1.0;
  double ediv = ekin/T->kT;
  double blow = -ediv;
  int starts_kin = 1;
  if (blow < bg[0]) {
    double c1 = bg[0] - (bg[1]-bg[0])*1e-6, c2 = bg[0] - fabs(bg[0])*1e-13, c3 = nextafter(bg[0], blow);
    blow = orc_min(c3, orc_min(c2, c1));
    starts_kin = 0;
  }
  /* activeGridRanges */
  unsigned short* rlo = (unsigned short*)malloc(sizeof(unsigned short)*nb*2);
  unsigned short* rup = rlo + nb;
  int nranges = 0, ibeta_low = 0;
  int itLow = 0, itUpp = na-1;
  for (int ib = 0; ib < nb; ++ib) {
    double alow = -1.0, aupp = -2.0;
    if (bg[ib] > -ediv) alpha_limits(ediv, bg[ib], &alow, &aupp);
    if (ag[na-1] <= alow || ag[0] >= aupp || aupp < alow) {
      if (nranges == 0) ibeta_low = ib + 1;
      else { rlo[nranges] = (unsigned short)na; rup[nranges] = (unsigned short)na; ++nranges; }
      continue;
    }
    while (ag[itLow] > alow && itLow > 0) --itLow;
    while (itLow < na-1 && ag[itLow+1] <= alow) ++itLow;
    if (itUpp < itLow) itUpp = itLow;
    while (ag[itUpp] < aupp && itUpp < na-1) ++itUpp;
    while (itUpp > 0 && ag[itUpp-1] >= aupp) --itUpp;
    rlo[nranges] = (unsigned short)itLow; rup[nranges] = (unsigned short)itUpp; ++nranges;
  }
  if (ibeta_low >= nb) { free(rlo); return 0; }   /* SABSamplerAtE_NoScatter */
  if (ibeta_low > 0 && blow < bg[ibeta_low-1]) { blow = bg[ibeta_low-1]; starts_kin = 0; }
  int nrel = nb - ibeta_low;
  double* vals = (double*)malloc(sizeof(double)*(nrel+1)*3);
  double* wts = vals + (nrel+1);
  double* cdf = wts + (nrel+1);
  orc_ainfo* infos = (orc_ainfo*)calloc(nrel, sizeof(orc_ainfo));
  int np = 0, ni = 0;
  double prev_b = blow, prev_xs = 0.0;
  vals[np] = prev_b; wts[np] = prev_xs; ++np;
  orc_ssum tot = {0, 0};
  int next_kin = starts_kin;
  int bad = 0;
  for (int k = 0; k < nrel; ++k) {
    double beta = bg[ibeta_low + k];
    if (beta == blow) { bad = 1; continue; }
    double alow, aupp;
    alpha_limits(ediv, beta, &alow, &aupp);
    double xs_here = 0.0;
    unsigned il = rlo[k], iu = rup[k];
    tailed tb; memset(&tb, 0, sizeof(tb));
    if (iu > il && aupp > alow) {
      size_t off = (size_t)na*(size_t)(k + ibeta_low);
      tb = tailed_breakdown(ag, na, T->sab + off, T->logsab + off, T->cumul + off, alow, aupp, il, iu);
      xs_here = tb.xs_front + tb.xs_back + tb.xs_middle;
    }
    orc_ainfo* info = &infos[ni++];
    if (xs_here > 0.0) {
      info->f_alpha = tb.fa; info->f_sval = tb.fs; info->f_logsval = tb.fl;
      info->b_alpha = tb.ba; info->b_sval = tb.bs; info->b_logsval = tb.bl;
      if (tb.narrow) info->prob_front = 1.0;
      else {
        info->prob_front = tb.xs_front/xs_here;
        info->prob_notback = 1.0 - tb.xs_back/xs_here;
        info->f_idx = (int)tb.imid_lo; info->b_idx = (int)tb.imid_up;
      }
    } else { info->prob_front = 2.0; info->f_alpha = alow; info->b_alpha = aupp; }
    if (next_kin) {
      next_kin = 0;
      double db = beta - prev_b;
      prev_b -= db*(1.0/3.0);
      vals[0] = prev_b;
    }
    orc_ssum_add(&tot, 0.5*(beta - prev_b)*(xs_here + prev_xs));
    prev_b = beta; prev_xs = xs_here;
    vals[np] = prev_b; wts[np] = prev_xs; ++np;
  }
  free(rlo);
  double xs_total = orc_ssum_get(&tot)*T->bound_xs/(4*ediv);
  if (!(xs_total >= 0.0)) xs_total = 0.0;
  ep->xs_check = xs_total;
  if (bad || xs_total == 0.0) { free(vals); free(infos); return bad ? -1 : 0; }
  /* PointwiseDist ctor, ref: src/utils/NCPointwiseDist.cc:32-74 */
  orc_ssum area = {0, 0};
  cdf[0] = 0.0;
  for (int i = 1; i < np; ++i) { orc_ssum_add(&area, (vals[i]-vals[i-1])*0.5*(wts[i]+wts[i-1])); cdf[i] = orc_ssum_get(&area); }
  double totarea = orc_ssum_get(&area);
  if (!(totarea > 0.0)) { free(vals); free(infos); return -2; }
  double nf = 1.0/totarea;
  for (int i = 0; i < np; ++i) { cdf[i] *= nf; wts[i] *= nf; }
  cdf[np-1] = 1.0;
  ep->npts = np; ep->ibeta_off = ibeta_low;
  ep->first_bin = starts_kin ? blow : 1.0;
  ep->x = vals; ep->pdf = wts; ep->cdf = cdf; ep->infos = infos;
  return 0;
}

/* SABData2DerivedDataFactory::actualCreate + SABIntegrator::Impl::doit loop, ref: NCSABIntegrator.cc:105-141,283-315 */
int orc_sab_build(orc_sab* T)
{
  size_t n = (size_t)T->nalpha*T->nbeta;
  T->logsab = (double*)malloc(sizeof(double)*n);
  T->cumul = (double*)calloc(n, sizeof(double));
  for (size_t i = 0; i < n; ++i) T->logsab[i] = T->sab[i] > 0.0 ? log(T->sab[i]) : -HUGE_VAL;
  for (int ib = 0; ib < T->nbeta; ++ib) {
    size_t off = (size_t)ib*T->nalpha;
    double cum = 0.0;
    for (int ai = 0; ai + 1 < T->nalpha; ++ai) {
      cum += integ_fast(T->alpha[ai], T->sab[off+ai], T->alpha[ai+1], T->sab[off+ai+1], T->logsab[off+ai], T->logsab[off+ai+1]);
      T->cumul[off+ai+1] = cum;
    }
  }
  T->ep = (orc_epoint*)calloc(T->negrid, sizeof(orc_epoint));
  for (int ie = 0; ie < T->negrid; ++ie)
    if (analyse_epoint(T, T->egrid[ie], &T->ep[ie]) < 0) return -1;
  return 0;
}
void orc_sab_free(orc_sab* T)
{
  if (T->ep) for (int i = 0; i < T->negrid; ++i) { free(T->ep[i].x); free(T->ep[i].infos); }
  free(T->ep); free(T->logsab); free(T->cumul);
  T->ep = 0; T->logsab = T->cumul = 0;
}

/* ------------------------------------------------------------------ SAB: sampling */
/* sampleLogLinDist_fast, ref: NCSABUtils.hh:282-303 */
static double loglin_fast(double a, double fa, double b, double fb, double r, double la, double lb)
{
  double df = fb - fa;
  if (fa*fb*df != 0.0) {
    double amb = a - b, l = lb - la;
    if (amb*l != 0.0) return amb*log(fa*exp(a*l/amb)/(fa + r*df))/l;
    df = 0.0;
  }
  if (!df) return a + r*(b-a);
  double x = (b-a)*sqrt(r);
  return fa ? b - x : a + x;
}
/* PointwiseDist::percentileWithIndex, ref: src/utils/NCPointwiseDist.cc:76-105 */
static double pwd_percentile(const orc_epoint* ep, double p, int* idx)
{
  int n = ep->npts;
  if (p == 1.) { *idx = n-2; return ep->x[n-1]; }
  int i = orc_lower_bound(ep->cdf, 0, n, p);
  if (i > n-1) i = n-1;
  if (i < 1) i = 1;
  double dx = ep->x[i] - ep->x[i-1], c = p - ep->cdf[i-1], a = ep->pdf[i-1], d = ep->pdf[i] - a, zdx;
  if (!a) zdx = d > 0.0 ? sqrt((2.0*c*dx)/d) : 0.5*dx;
  else {
    double e = d*c/(dx*a*a);
    if (fabs(e) > 1e-7) zdx = (sqrt(1.0 + 2.0*e) - 1.0)*dx*a/d;
    else zdx = (1 + 0.5*e*(e - 1.0))*c/a;
  }
  *idx = i-1;
  return orc_clamp(ep->x[i-1] + zdx, ep->x[i-1], ep->x[i]);
}
/* SABSamplerAtE_Alg1::sampleAlpha, ref: src/sab/NCSABSamplerModels.cc:157-233 */
static double sab_sample_alpha(const orc_sab* T, const orc_epoint* ep, int ibeta, double r)
{
  const orc_ainfo* f = &ep->infos[ibeta - ep->ibeta_off];
  size_t off = (size_t)ibeta*T->nalpha;
  const double *cum = T->cumul + off, *sab = T->sab + off, *ls = T->logsab + off, *ag = T->alpha;
  if (r <= f->prob_front) {
    if (f->prob_front == 2.0) return f->f_alpha + r*(f->b_alpha - f->f_alpha);
    if (f->prob_front == 1.0) return loglin_fast(f->f_alpha, f->f_sval, f->b_alpha, f->b_sval, r, f->f_logsval, f->b_logsval);
    double p2 = orc_clamp(r/f->prob_front, DBL_MIN, 1.0);
    return loglin_fast(f->f_alpha, f->f_sval, ag[f->f_idx], sab[f->f_idx], p2, f->f_logsval, ls[f->f_idx]);
  } else if (r <= f->prob_notback) {
    double p2 = orc_clamp((r - f->prob_front)/(f->prob_notback - f->prob_front), 0.0, 1.0);
    int il = f->f_idx, iu = f->b_idx;
    double area = cum[il] + p2*(cum[iu] - cum[il]);
    int is = orc_upper_bound(cum, il, iu+1, area);
    if (is > iu) return ag[iu];
    if (is <= il) return ag[il];
    int a0 = is-1, a1 = is;
    double binArea = cum[a1] - cum[a0];
    double rr = orc_clamp((area - cum[a0])/binArea, DBL_MIN, 1.0);
    return loglin_fast(ag[a0], sab[a0], ag[a1], sab[a1], rr, ls[a0], ls[a1]);
  }
  double p2 = orc_clamp((r - f->prob_notback)/(1.0 - f->prob_notback), DBL_MIN, 1.0);
  return loglin_fast(ag[f->b_idx], sab[f->b_idx], f->b_alpha, f->b_sval, p2, ls[f->b_idx], f->b_logsval);
}
/* SABSamplerAtE_Alg1::sampleAlphaBeta, ref: NCSABSamplerModels.cc:48-155 */
static void sab_sample_at_e(const orc_sab* T, const orc_epoint* ep, double ediv, orc_rng* rng, double* alpha, double* beta_out, int* err)
{
  if (ep->npts == 0) { *alpha = 0.0; *beta_out = 0.0; return; }   /* SABSamplerAtE_NoScatter */
  const double* bg = T->beta;
  for (int loop = 0; loop < 100; ++loop) {
    int ib;
    double beta = pwd_percentile(ep, orc_rand(rng), &ib);
    if (ib == 0 && ep->first_bin <= 0.0) {
      double b0 = ep->first_bin, b1 = ep->x[1];
      if (b1 < -ediv) continue;
      double db = b1 - b0, aval = 0.0, lo, up;
      for (int iii = 0; iii < 30; ++iii) {
        beta = orc_max(ep->first_bin, b0 + db*orc_rand(rng));
        if (beta < -ediv) break;
        aval = sab_sample_alpha(T, ep, ep->ibeta_off, orc_rand(rng));
        alpha_limits(-ep->first_bin, beta, &lo, &up);
        if (orc_in(lo, up, aval)) break;
        if (iii == 29) { aval = 0.5*(lo + up); break; }
      }
      if (beta < -ediv) continue;
      alpha_limits(ediv, beta, &lo, &up);
      if (orc_in(lo, up, aval)) { *alpha = aval; *beta_out = beta; return; }
      continue;
    }
    if (beta <= orc_max(-ediv, bg[0])) continue;
    double r = orc_rand(rng);
    int ibeta = ep->ibeta_off + ib;
    double bl = bg[ibeta-1], al = sab_sample_alpha(T, ep, ibeta-1, r);
    double bh = bg[ibeta], ah = sab_sample_alpha(T, ep, ibeta, r);
    double a = al + (ah - al)*(beta - bl)/(bh - bl), lo, up;
    alpha_limits(ediv, beta, &lo, &up);
    if (orc_in(lo, up, a)) { *alpha = a; *beta_out = beta; return; }
  }
  *err |= ORC_ERR_INNER; *alpha = -1.0; *beta_out = 0.0;
}
/* SABSampler::sampleHighE, ref: src/sab/NCSABSampler.cc:59-156; returns 1 when (alpha,beta) is final */
static int sab_high_e(const orc_sab* T, double ekin, orc_rng* rng, double* alpha, double* beta, int* err)
{
  double emax = T->egrid[T->negrid-1];
  double xe = ekin*fg_xs(&T->ext, ekin);
  double Pin = T->k1/((T->k1 - T->k2) + xe);
  double Pext = T->k2/xe;
  double Pdis = (Pext >= Pin ? (1.0 - Pin/Pext) : 0.0);
  if (Pdis > 0.95) { *err |= ORC_ERR_DISCARD; *alpha = -1.0; *beta = 0.0; return 1; }
  if (Pext < Pin) {
    double aa = 1.0 - Pext;
    double Pextra = aa > 1e-10 ? (Pin - Pext)/aa : 1.0;
    if (orc_rand(rng) < Pextra) return 0;
  }
  double emax_div = emax/T->kT;
  fgsampler s; fgs_init(&s, ekin, T->ext.kT, T->ext.mass_amu);
  for (;;) {
    double lo, up;
    fgs_alpha_beta(&s, rng, alpha, beta);
    if (*beta <= -emax_div) return 1;
    alpha_limits(emax_div, *beta, &lo, &up);
    if (!orc_in(lo, up, *alpha)) return 1;
    if (Pdis && orc_rand(rng) < Pdis) continue;
    return 0;
  }
}
/* SABSampler::sampleAlphaBeta + sampleDeltaEMu + SABScatter::sampleScatterIsotropic,
 * ref: NCSABSampler.cc:158-236, src/sabscatter/NCSABScatter.cc:93-100 */
static void sab_sample(const orc_sab* T, double ekin_in, orc_rng* rng, double* eout, double* mu, int* err)
{
  double ekin = ekin_in, alpha = 0.0, beta = 0.0;
  int n = T->negrid, iu = orc_upper_bound(T->egrid, 0, n, ekin), isamp, ultra = 0, have = 0;
  if (iu == n) {
    if (sab_high_e(T, ekin, rng, &alpha, &beta, err)) have = 1;
    else { ekin = T->egrid[n-1]; isamp = n-1; }
  } else if (iu == 0) { isamp = 0; ultra = (ekin < T->egrid[0]); }
  else {
    if (T->egrid_margin > 1.0) while (iu + 1 != n && ekin*T->egrid_margin > T->egrid[iu]) ++iu;
    isamp = iu;
  }
  if (*err & ORC_ERR_DISCARD) { *eout = -1.0; *mu = -999.0; return; }
  if (!have) {
    double ediv = ekin/T->kT;
    double sdiv = ultra ? T->egrid[0]/T->kT : ediv;
    int ok = 0;
    for (int loop = 0; loop < 100 && !ok; ++loop) {
      double lo, up;
      sab_sample_at_e(T, &T->ep[isamp], sdiv, rng, &alpha, &beta, err);
      if (*err & ORC_ERR_INNER) { *eout = -1.0; *mu = -999.0; return; }
      if (beta < -ediv) continue;
      alpha_limits(ediv, beta, &lo, &up);
      if (orc_in(lo, up, alpha)) { ok = 1; break; }
      if (ultra) { alpha = lo + orc_rand(rng)*(up - lo); ok = 1; break; }
    }
    if (!ok) { *err |= ORC_ERR_OUTER; *eout = -1.0; *mu = -999.0; return; }
  }
  double dE;
  if (mu_iso_at_beta(beta, ekin_in/T->kT)) { dE = beta*T->kT; *mu = orc_rand(rng)*2.0 - 1.0; }
  else {
    ab_to_demu(alpha, beta, ekin_in, T->kT, &dE, mu, err);
    if (*err & ORC_ERR_KIN) { *eout = -1.0; *mu = -999.0; return; }
  }
  *eout = orc_max(0.0, ekin_in + dE);
}

/* ------------------------------------------------------------------ composition */
double orc_comp_xs_iso(const orc_material* M, int i, double ekin, int* aux)
{
  const orc_comp* c = &M->comp[i];
  *aux = -1;
  switch (c->kind) {
  case NCB_KIND_POWDERBRAGG: return pb_xs(&M->pb[c->idx], ekin, aux);
  case NCB_KIND_ELINC: return elinc_xs(&M->elinc[c->idx], ekin, 0);
  case NCB_KIND_SAB: return sab_xs(&M->sab[c->idx], ekin);
  case NCB_KIND_FREEGAS: return fg_xs(&M->fg[c->idx], ekin);
  default: return 0.0;
  }
}
/* ProcComposition::Impl::updateCacheIsotropic / crossSectionIsotropic, ref: src/interfaces/NCProcImpl.cc:166-204,353-362 */
double orc_xs_iso(const orc_material* M, double ekin, double* cumul, int* aux)
{
  if (!orc_domain_contains(M->dom_lo, M->dom_hi, ekin)) return 0.0;
  double tot = 0.0;
  for (int i = 0; i < M->ncomp; ++i) {
    int a = -1;
    double xs = orc_domain_contains(M->comp[i].dom_lo, M->comp[i].dom_hi, ekin) ? orc_comp_xs_iso(M, i, ekin, &a) : 0.0;
    tot += M->comp[i].scale*xs;
    if (cumul) cumul[i] = tot;
    if (aux) aux[i] = a;
  }
  return tot;
}
void orc_comp_sample_iso(const orc_material* M, int i, int aux, double ekin, orc_rng* rng, double* eout, double* mu, int* err)
{
  const orc_comp* c = &M->comp[i];
  switch (c->kind) {
  case NCB_KIND_POWDERBRAGG: pb_sample(&M->pb[c->idx], ekin, aux, rng, eout, mu); return;
  case NCB_KIND_ELINC: *eout = ekin; *mu = elinc_sample_mu(&M->elinc[c->idx], ekin, rng); return;   /* NCElIncScatter.cc:198-204 */
  case NCB_KIND_SAB: sab_sample(&M->sab[c->idx], ekin, rng, eout, mu, err); return;
  case NCB_KIND_FREEGAS: {                                                                        /* NCFreeGas.cc:70-75 */
    fgsampler s; double dE;
    fgs_init(&s, ekin, M->fg[c->idx].kT, M->fg[c->idx].mass_amu);
    fgs_demu(&s, rng, &dE, mu, err);
    *eout = orc_max(0.0, ekin + dE);
    return;
  }
  default: *eout = ekin; *mu = 1.0; return;
  }
}
/* ProcComposition::sampleScatterIsotropic, ref: NCProcImpl.cc:379-389 */
void orc_sample_iso(const orc_material* M, double ekin, orc_rng* rng, double* eout, double* mu, int* err)
{
  if (!orc_domain_contains(M->dom_lo, M->dom_hi, ekin)) { *eout = ekin; *mu = 1.0; return; }
  double cumul[ORC_MAXCOMP]; int aux[ORC_MAXCOMP];
  orc_xs_iso(M, ekin, cumul, aux);
  int ich = (M->ncomp == 1 ? 0 : orc_pick(orc_rand(rng), cumul, M->ncomp));
  orc_comp_sample_iso(M, ich, aux[ich], ekin, rng, eout, mu, err);
}
