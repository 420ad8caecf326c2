/* oracle/oracle_mmc.c -- TEST INFRASTRUCTURE (see oracle_common.h).  CPU restatement of the reference's MiniMC
 * standard engine for one volume of one material, one neutron HISTORY at a time (the reference works on baskets of
 * 4096 neutrons, the product on whole populations; the per-neutron arithmetic is the same):
 *   step        ncrystal_core/src/minimc/NCMMC_SimEngine.cc:168-455
 *   helpers     src/minimc/NCMMC_Utils.cc (calcProbTransm, propagateAndAttenuate, sampleRandDists, distToSlab*)
 *   geometry    src/minimc/NCMMC_Sphere.hh, NCMMC_Slab.hh, NCMMC_Box.hh, NCMMC_Cyl.hh
 *   sources     src/minimc/NCMMC_Source.cc (constant :545-640, circular :642-800, energies :205-300)
 *   entry       src/minimc/NCMMC_BasketSrcFiller.hh:84-160
 *   tallies     src/minimc/NCMMC_StdTallies.cc:240-640, include/NCrystal/internal/utils/NCHists.hh:378-396
 * Random numbers: the product's per-(neutron, step) Philox streams (philox_ref.h; ncb_mmc.cuh header).  The
 * reference MiniMC itself draws from per-thread streams in basket order, so it can only be compared statistically
 * (its own acceptance test: chi-square of the exit-angle histogram, tests/pypath/NCTestUtils/minimc_ref.py); the
 * product is compared with THIS restatement history by history. */
#include "oracle_common.h"
#include <float.h>

#define MMC_SID_SRC  0x4D4D0000u
#define MMC_SID_BASE 0x4D4D0010u

typedef struct {
  int geom_kind;            /* 1 sphere (ga=r), 2 slab (gc=dz), 3 box (ga,gb,gc = dx,dy,dz), 4 cyl (ga=r, gb=dy or 0) */
  double ga, gb, gc;
  int src_kind;             /* 1 constant, 2 circular, 3 isotropic (radius = signed r: start at pos - r*dir) */
  double pos[3], dir[3];    /* dir: any non-null vector */
  double radius;
  int emode;                /* 0 fixed ekin e0, 1 uniform ekin in [e0,e1], 2 uniform wavelength in [e0,e1],
                               3 log-normal ekin (mean e0, rms e1), 4 log-normal wavelength, 5 Maxwell at e0 kelvin */
  double e0, e1;
  double weight;
  double roul_psurv, roul_wthr; int roul_nscat;
  int nscatlimit;           /* -1 none */
  int ignore_miss, include_abs;
  uint64_t seed;
} orc_mmc_cfg;

/* ---- geometry ---- */
static double slab_exit(double x, double ux, double d)
{
  if (ux > 0.0) return (d - x) / ux;
  if (ux < 0.0) return -(d + x) / ux;
  return INFINITY;
}
static double slab_entry(double x, double ux, double d)
{
  double a = fabs(x) - d, xu = x * ux;
  if (a <= 0) { if (a) return 0.0; return xu > 0.0 ? -1.0 : 0.0; }
  if (xu >= 0.0) return -1.0;
  return a / fabs(ux);
}
static void cyl_pars(const orc_mmc_cfg* c, const double* p, const double* u, double* twoA, double* B, double* C, double* D)
{
  double rsq = c->ga * c->ga;
  *twoA = u[0]*u[0]; *twoA += u[2]*u[2]; *twoA *= 2;
  *B = p[0]*u[0]; *B += p[2]*u[2]; *B *= 2;
  *C = p[0]*p[0]; *C += p[2]*p[2]; *C -= rsq;
  *D = (*twoA) * (*C); *D *= -2.0; *D += (*B) * (*B);
}
static double dist_exit(const orc_mmc_cfg* c, const double* p, const double* u)
{
  if (c->geom_kind == 1) {
    double t = -p[0]*p[0]; t -= p[1]*p[1]; t -= p[2]*p[2]; t += c->ga * c->ga;
    double pd = p[0]*u[0]; pd += p[1]*u[1]; pd += p[2]*u[2];
    t += pd * pd;
    t = sqrt(fmax(0.0, t));
    t -= pd;
    return t > 0.0 ? t : 0.0;
  }
  if (c->geom_kind == 2) return slab_exit(p[2], u[2], c->gc);
  if (c->geom_kind == 3) {
    double t = slab_exit(p[0], u[0], c->ga), t2 = slab_exit(p[1], u[1], c->gb);
    if (t2 < t) t = t2;
    t2 = slab_exit(p[2], u[2], c->gc);
    return t2 < t ? t2 : t;
  }
  double twoA, B, C, D;
  cyl_pars(c, p, u, &twoA, &B, &C, &D);
  double t = sqrt(fabs(D));
  t -= B;
  if (!(t > 0.0)) t = 0.0;
  if (C > 0.0) C = 0.0;
  int done = 0;
  if (twoA * C == 0.0) {
    if (!twoA) { t = INFINITY; done = 1; }
    else if (B >= 0.0) { t = 0.0; done = 1; }
  }
  if (!done) t /= twoA;
  if (c->gb) { double ts = slab_exit(p[1], u[1], c->gb); if (ts < t) t = ts; }
  return t;
}
static double dist_entry(const orc_mmc_cfg* c, const double* p, const double* u)
{
  if (c->geom_kind == 1) {
    double pdotu = p[0]*u[0]; pdotu += p[1]*u[1]; pdotu += p[2]*u[2];
    double psq = p[0]*p[0]; psq += p[1]*p[1]; psq += p[2]*p[2]; psq -= c->ga * c->ga;
    if (psq <= 0.0) return (psq < pdotu ? psq : pdotu) < 0.0 ? 0.0 : -1.0;
    double D = pdotu * pdotu - psq;
    if (D < 0) return -1.0;
    double t = -(sqrt(D) + pdotu);
    return t >= 0.0 ? t : -1.0;
  }
  if (c->geom_kind == 2) return slab_entry(p[2], u[2], c->gc);
  if (c->geom_kind == 3) {
    const double half[3] = { c->ga, c->gb, c->gc };
    double t1 = -INFINITY, t2 = INFINITY;
    for (int k = 0; k < 3; ++k) {
      if (u[k]) {
        double inv = 1.0 / u[k], q1 = (half[k] - p[k]) * inv, q2 = (-half[k] - p[k]) * inv;
        double lo = q1 < q2 ? q1 : q2, hi = q1 < q2 ? q2 : q1;
        if (lo > t1) t1 = lo;
        if (hi < t2) t2 = hi;
      } else if (fabs(p[k]) > half[k]) return -1.0;
    }
    if (t1 >= t2 || t2 <= 0.0) return -1.0;
    return t1 > 0.0 ? t1 : 0.0;
  }
  double twoA, B, C, D;
  cyl_pars(c, p, u, &twoA, &B, &C, &D);
  if (!c->gb) {
    if (twoA == 0.0) return C <= 0.0 ? 0.0 : -1.0;
    if (C <= 0.0) return (C == 0.0 && B >= 0.0) ? -1.0 : 0.0;
    if (D <= 0) return -1.0;
    double s = sqrt(D), tmin = -B - s, tmax = -B + s;
    if (tmax < 0.0) return -1.0;
    if (tmin > 0.0) return tmin / twoA;
    return B < 0.0 ? 0.0 : -1.0;
  }
  double tmin, tmax;
  if (twoA == 0.0) {
    if (C <= 0) { tmin = -INFINITY; tmax = INFINITY; } else return -1.0;
  } else {
    if (D <= 0) return -1.0;
    double s = sqrt(D), inv = 1.0 / twoA;
    tmin = fmax(0.0, (-B - s) * inv); tmax = fmax(0.0, (-B + s) * inv);
  }
  if (fabs(p[1]) == c->gb && p[1] * u[1] > 0.0) return -1.0;
  if (u[1] == 0.0) { if (fabs(p[1]) > c->gb) return -1.0; }
  else {
    double t1 = -(p[1] + c->gb) / u[1], t2 = (c->gb - p[1]) / u[1];
    double lo = t1 < t2 ? t1 : t2, hi = t1 < t2 ? t2 : t1;
    tmin = fmax(0.0, lo > tmin ? lo : tmin);
    tmax = fmax(0.0, hi < tmax ? hi : tmax);
  }
  if (!(tmax > tmin)) return -1.0;
  return tmin;
}
static int point_inside(const orc_mmc_cfg* c, const double* p)
{
  if (c->geom_kind == 1) return p[0]*p[0] + p[1]*p[1] + p[2]*p[2] <= c->ga * c->ga;
  if (c->geom_kind == 2) return fabs(p[2]) <= c->gc;
  if (c->geom_kind == 3) return fabs(p[0]) <= c->ga && fabs(p[1]) <= c->gb && fabs(p[2]) <= c->gc;
  return p[0]*p[0] + p[2]*p[2] <= c->ga * c->ga && (c->gb == 0 || fabs(p[1]) <= c->gb);
}

/* ---- transmission ---- */
static double prob_transm(int has_xs, double xs, double dist, int unb)
{
  if (!has_xs) return unb ? (isinf(dist) ? 0.0 : 1.0) : 1.0;
  double t = xs < DBL_MAX ? xs : DBL_MAX;
  if (unb) t *= (t ? dist : 0.0); else t *= dist;
  t = exp(-t);
  if (unb) t *= (isinf(dist) ? 0.0 : 1.0);
  return t;
}

/* ---- tallies ---- */
enum { T_THETA, T_MU, T_NSCAT, T_NSCAT_UW, T_W, T_E, T_L, T_DE, T_Q };
#define NCLASS 5
#define NSTAT 5
typedef struct { int type, nbins; double xmin, xmax, invdelta; double* g; } hist_t;
typedef struct { int nh; hist_t h[9]; double dir0[3]; int has_dir0; int has_e0; double e0; double tallied_w; uint64_t tallied_n; } tally_t;

/* randNorm, ref: src/utils/NCRandUtils.cc:113-137.  (Its `continue` inside the do-while jumps to the loop condition,
 * so the quick-reject line before it has no effect on the outcome and is left out.) */
static double rand_norm(orc_rng* r)
{
  double g, g2, u, v, invu;
  do {
    u = orc_rand(r);
    invu = 1.0 / u;
    v = orc_rand(r);
    g = 1.71552776992141354 * (v - 0.5) * invu;
    g2 = g * g;
    if (g2 <= 5.0 - 5.13610166675096558 * u) break;
  } while (g2 >= -4.0 * log(u));
  return g;
}

static int value_to_bin(const hist_t* h, double v)
{
  if (v < h->xmin) return 0;
  if (v >= h->xmax) return v == h->xmax ? h->nbins : h->nbins + 1;
  uint64_t k = (uint64_t)(h->invdelta * (v - h->xmin));
  return 1 + (int)(k < (uint64_t)h->nbins ? k : (uint64_t)h->nbins);
}
static void tally_record(tally_t* T, const double* u, double ekin, double w, int nscat, int ninel, double e_init,
                         const double* u_init)
{
  static const double kToDeg = 57.2957795130823208767981548141051703324054725;
  int cls = nscat > 1 ? (ninel ? 4 : 3) : (nscat == 1 ? (ninel ? 2 : 1) : 0);
  T->tallied_w += w; T->tallied_n += 1;
  double mu;
  if (!T->has_dir0) { mu = u_init[0]*u[0]; mu += u_init[1]*u[1]; mu += u_init[2]*u[2]; }
  else if (T->dir0[2] == 1) mu = u[2];
  else { mu = T->dir0[0]*u[0]; mu += T->dir0[1]*u[1]; mu += T->dir0[2]*u[2]; }
  if (mu < -1.0) mu = -1.0;
  if (mu > 1.0) mu = 1.0;
  const double Ei = T->has_e0 ? T->e0 : e_init;
  for (int ih = 0; ih < T->nh; ++ih) {
    hist_t* h = &T->h[ih];
    double v, wgt = w;
    switch (h->type) {
    case T_THETA: v = acos(mu) * kToDeg; break;
    case T_MU: v = mu; break;
    case T_NSCAT: v = nscat; break;
    case T_NSCAT_UW: v = nscat; wgt = 1.0; break;
    case T_W: v = w; wgt = 1.0; break;
    case T_E: v = ekin; break;
    case T_L: v = 1.0 / fmax(4.9406564584124654e-324, ekin); v = sqrt(v); v *= 0.2860143520967626; break;
    case T_DE: v = Ei - ekin; break;
    default: v = ekin * Ei; v = sqrt(v); v *= mu; v *= -2.0; v += Ei; v += ekin;
             v *= 39.4784176043574344753379639995046045412547976 * 12.22430978582345950656; v = sqrt(fmax(0.0, v)); break;
    }
    if (!(wgt > 0.0)) continue;
    const int nb2 = h->nbins + 2, bin = value_to_bin(h, v);
    h->g[cls * nb2 + bin] += wgt;
    h->g[NCLASS * nb2 + cls * nb2 + bin] += wgt * wgt;
    double* s = h->g + 2 * NCLASS * nb2 + cls * NSTAT;
    if (v < s[3]) s[3] = v;
    if (v > s[4]) s[4] = v;
    s[0] += wgt; s[1] += wgt * v; s[2] += wgt * v * v;
  }
}

/* Runs source neutrons [first, first+count).  Histogram i occupies NCLASS*(2*(nbins+2)+NSTAT) doubles of `out`
 * (per class: contents incl. under/overflow, squared weights, then sumw, sumwx, sumwx2, min, max), concatenated.
 * meta: [0] missed count, [1] missed weight, [2] tallied count, [3] tallied weight, [4] total steps. */
int orc_minimc_run(void* vm, const orc_mmc_cfg* c, uint64_t first, uint64_t count, int ntally, const int* types,
                   const int* nbins, const double* xmin, const double* xmax, double* out, double* meta)
{
  const orc_material* M = (const orc_material*)vm;
  const ncb_header_t* hdr = (const ncb_header_t*)M->blob;
  const double macro = 100.0 * hdr->numdens;             /* Utils::macroXSFactor */
  const double abs_c = (c->include_abs && hdr->abs_c > 0.0) ? hdr->abs_c : 0.0;
  const int unb = (c->geom_kind == 2) || (c->geom_kind == 4 && c->gb == 0.0);
  tally_t T; memset(&T, 0, sizeof(T));
  T.nh = ntally;
  size_t off = 0;
  for (int i = 0; i < ntally; ++i) {
    hist_t* h = &T.h[i];
    h->type = types[i]; h->nbins = nbins[i]; h->xmin = xmin[i]; h->xmax = xmax[i];
    h->invdelta = 1.0 / ((xmax[i] - xmin[i]) / nbins[i]);
    h->g = out + off;
    size_t nd = (size_t)NCLASS * (2 * (nbins[i] + 2) + NSTAT);
    memset(h->g, 0, nd * sizeof(double));
    for (int k = 0; k < NCLASS; ++k) {   /* running min/max of the filled values: empty = (+inf,-inf) */
      h->g[2 * NCLASS * (nbins[i] + 2) + k * NSTAT + 3] = INFINITY;
      h->g[2 * NCLASS * (nbins[i] + 2) + k * NSTAT + 4] = -INFINITY;
    }
    off += nd;
  }
  double dir[3];
  { double m = 1.0 / sqrt(c->dir[0]*c->dir[0] + c->dir[1]*c->dir[1] + c->dir[2]*c->dir[2]);
    for (int k = 0; k < 3; ++k) dir[k] = c->dir[k] * m; }
  memcpy(T.dir0, dir, sizeof(dir));
  T.has_e0 = (c->emode == 0); T.e0 = c->e0;
  T.has_dir0 = (c->src_kind != 3);
  /* log-normal / Maxwell parameters (NCMMC_ParseCfg.hh:414-427, :372) */
  double mu_n = 0, sigma_n = 0, halfkT = 0;
  if (c->emode == 3 || c->emode == 4) {
    const double tmp = 1 + (c->e1 * c->e1) / (c->e0 * c->e0);
    mu_n = log(c->e0 / sqrt(tmp)); sigma_n = sqrt(log(tmp));
  }
  if (c->emode == 5) halfkT = 8.6173303e-5 * c->e0 * 0.5;
  double va[3] = {0,0,0}, vb[3] = {0,0,0};
  if (c->src_kind == 2 && c->radius > 0.0) {
    double a[3] = {1,0,0};
    if (a[0]*dir[0] + a[1]*dir[1] + a[2]*dir[2] > 0.8) { a[0] = 0; a[1] = 1; }
    if (a[0]*dir[0] + a[1]*dir[1] + a[2]*dir[2] > 0.8) { a[1] = 0; a[2] = 1; }
    double t[3] = { dir[1]*a[2] - dir[2]*a[1], dir[2]*a[0] - dir[0]*a[2], dir[0]*a[1] - dir[1]*a[0] };
    double g = 1.0 / sqrt(t[0]*t[0] + t[1]*t[1] + t[2]*t[2]);
    for (int k = 0; k < 3; ++k) va[k] = t[k] * g;
    double s[3] = { dir[1]*va[2] - dir[2]*va[1], dir[2]*va[0] - dir[0]*va[2], dir[0]*va[1] - dir[1]*va[0] };
    g = 1.0 / sqrt(s[0]*s[0] + s[1]*s[1] + s[2]*s[2]);
    for (int k = 0; k < 3; ++k) { vb[k] = s[k] * g * c->radius; va[k] *= c->radius; }
  }
  const double minus_r = (c->src_kind == 3 && c->radius) ? -c->radius : 0.0;
  /* SourceIsotropic::particlesMightBeOutside (NCMMC_Source.cc:503-506) looks at (x,y,z) only, whatever r is */
  const int may_be_outside = (c->src_kind == 2 && c->radius > 0.0) ? 1 : !point_inside(c, c->pos);
  double miss_n = 0, miss_w = 0, nsteps = 0;
  int errs = 0;
  for (uint64_t id = first; id < first + count; ++id) {
    orc_rng r; ncb_stream_init_sid(&r, c->seed, id, MMC_SID_SRC);
    double p[3] = { c->pos[0], c->pos[1], c->pos[2] }, u[3] = { dir[0], dir[1], dir[2] };
    if (c->src_kind == 2 && c->radius > 0.0) {
      double a, b;
      do { a = -1.0 + orc_rand(&r) * 2.0; b = -1.0 + orc_rand(&r) * 2.0; } while (a*a + b*b > 1.0);
      for (int k = 0; k < 3; ++k) p[k] = c->pos[k] + va[k] * a + vb[k] * b;
    }
    if (c->src_kind == 3) {
      /* randIsotropicDirection, src/utils/NCRandUtils.cc:25-49 */
      double x0, x1, ss;
      do { x0 = 2.0 * orc_rand(&r) - 1.0; x1 = 2.0 * orc_rand(&r) - 1.0; ss = x0*x0 + x1*x1; } while (ss >= 1.0);
      const double t = 2.0 * sqrt(1.0 - ss);
      u[0] = x0 * t; u[1] = x1 * t; u[2] = 1.0 - 2.0 * ss;
    }
    double w = c->weight, ekin;
    if (c->emode == 0) ekin = c->e0;
    else if (c->emode == 5) {
      double v = rand_norm(&r); v *= v;
      double g = rand_norm(&r); v += g * g;
      g = rand_norm(&r); v += g * g;
      ekin = v * halfkT;
    } else if (c->emode == 3 || c->emode == 4) {
      double v = rand_norm(&r);
      v *= sigma_n; v += mu_n;
      v = exp(v);
      if (c->emode == 4) { v *= v; v = 1.0 / fmax(4.9406564584124654e-324, v); v *= 0.081804209605330899; }
      ekin = v;
    } else {
      double v = orc_rand(&r) * (c->e1 - c->e0);
      v += c->e0;
      if (v > c->e1) v = c->e1;
      if (c->emode == 2) { v *= v; v = 1.0 / fmax(4.9406564584124654e-324, v); v *= 0.081804209605330899; }
      ekin = v;
    }
    if (minus_r != 0.0) for (int k = 0; k < 3; ++k) p[k] += u[k] * minus_r;
    const double e_init = ekin;
    const double u_init[3] = { u[0], u[1], u[2] };
    int nscat = 0, ninel = 0;
    if (may_be_outside) {
      double d = dist_entry(c, p, u);
      if (d < 0.0) {
        miss_n += 1; miss_w += w;
        if (!c->ignore_miss) tally_record(&T, u, ekin, w, -1, 0, e_init, u_init);
        continue;
      }
      for (int k = 0; k < 3; ++k) p[k] += d * u[k];
    }
    for (uint32_t step = 0; ; ++step) {
      nsteps += 1;
      orc_rng rs; ncb_stream_init_sid(&rs, c->seed, id, MMC_SID_BASE + 2u * step);
      const double d_exit = dist_exit(c, p, u);
      double xs_a = 0.0;
      if (abs_c > 0.0) { double sq = sqrt(ekin); xs_a = (sq ? abs_c / sq : INFINITY) * macro; }
      orc_vec dv = { u[0], u[1], u[2] };
      double xs_s = macro * orc_xs(M, ekin, dv, 0, 0, 0);
      if (c->nscatlimit >= 0 && nscat >= c->nscatlimit) xs_s = 0.0;
      const double ptransm = prob_transm(1, xs_s, d_exit, unb);
      const double uni = orc_rand(&rs);
      double d_scat;
      if (!xs_s) d_scat = INFINITY;
      else if (isinf(d_exit)) d_scat = log(uni) / (-xs_s);
      else { double c1 = -1.0 / xs_s, c2 = expm1(-xs_s * (d_exit - 0.0)); d_scat = 0.0 + c1 * log(1.0 + uni * c2); }
      /* transmitted part */
      double wt = w;
      wt *= prob_transm(abs_c > 0.0, xs_a, d_exit, unb);
      wt *= ptransm;
      tally_record(&T, u, ekin, wt, nscat, ninel, e_init, u_init);
      /* scattered part */
      if (w == 0.0 || isinf(d_scat) || !(xs_s > 0.0)) break;
      double rfact = 1.0;
      if (nscat >= c->roul_nscat && w < c->roul_wthr) {
        if (orc_rand(&rs) > c->roul_psurv) break;
        rfact = 1.0 / c->roul_psurv;
      }
      const double wred = exp(-xs_a * d_scat);
      if (!(wred > 0.0)) break;
      w *= rfact;
      for (int k = 0; k < 3; ++k) p[k] += d_scat * u[k];
      w *= wred;
      orc_rng rq; ncb_stream_init_sid(&rq, c->seed, id, MMC_SID_BASE + 2u * step + 1u);
      double eout; orc_vec o; int err = 0;
      orc_sample(M, ekin, dv, &rq, &eout, &o, &err);
      errs |= err;
      u[0] = o.x; u[1] = o.y; u[2] = o.z;
      const int was_elastic = (ekin == eout);
      ekin = eout;
      ++nscat;
      if (!was_elastic) ++ninel;
      w *= (1.0 - ptransm);
      if (step > 100000u) return -1;
    }
  }
  meta[0] = miss_n; meta[1] = miss_w; meta[2] = (double)T.tallied_n; meta[3] = T.tallied_w; meta[4] = nsteps;
  return errs;
}
