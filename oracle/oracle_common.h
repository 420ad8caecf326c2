/* oracle/oracle_common.h -- TEST INFRASTRUCTURE.  Plain-C CPU restatement of the reference's
 * algorithm for the hot path (batched cross sections + scatter sampling).  Used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER of the CUDA path; it is
 * never linked, imported or executed by the product (ncrystal_b200/).
 *
 * Pinning: tests/test_cpu_oracle_port.py checks this restatement against the golden vectors
 * generated from the unmodified reference (tests/golden/, made by tests/golden/make_golden.py with
 * oracle/_ref) -- cross sections and replayed scatter outcomes -- so parity is PINNED.
 *
 * Every function cites the reference routine it follows (paths relative to
 * /root/reference/ncrystal_core).  Input: a compiled material (ncrystal_b200/csrc/ncb_blob.h).
 * Unlike the product (row-parallel table builder, split kernels) everything here is written the
 * way the reference runs it: sequentially, one neutron and one energy point at a time.
 */
#ifndef NCB_ORACLE_COMMON_H
#define NCB_ORACLE_COMMON_H
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ncb_blob.h"
#include "philox_ref.h"

#define ORC_MAXCOMP 8
#define ORC_MAXEL 16

/* constants, ref: include/NCrystal/core/NCDefs.hh:79-123,834-868 */
#define ORC_BOLTZMANN 8.6173303e-5
#define ORC_NEUTRON_MASS_AMU 1.00866491588
#define ORC_PI 3.1415926535897932384626433832795028841971694
#define ORC_PISQ 9.86960440108935861883449099987615113531369941
#define ORC_INVSQRTPI 0.564189583547756286948079451560772585844050629
#define ORC_EKIN2WLSQINV 12.22430978582345950656
#define ORC_WL2EKIN 0.081804209605330899

typedef ncb_stream_t orc_rng;
static inline double orc_rand(orc_rng* r) { return ncb_stream_next(r); }

static inline double orc_min(double a, double b) { return a < b ? a : b; }            /* ncmin */
static inline double orc_max(double a, double b) { return a > b ? a : b; }            /* ncmax */
static inline double orc_clamp(double v, double lo, double hi) { return orc_min(orc_max(v, lo), hi); } /* ncclamp */
static inline int orc_in(double a, double b, double x) { return (a <= x) && (x <= b); } /* valueInInterval */

/* std::upper_bound / std::lower_bound on [lo,hi) of a sorted array; return index */
static inline int orc_upper_bound(const double* a, int lo, int hi, double v)
{
  int len = hi - lo;
  while (len > 0) { int half = len >> 1; int mid = lo + half; if (v < a[mid]) len = half; else { lo = mid + 1; len = len - half - 1; } }
  return lo;
}
static inline int orc_lower_bound(const double* a, int lo, int hi, double v)
{
  int len = hi - lo;
  while (len > 0) { int half = len >> 1; int mid = lo + half; if (a[mid] < v) { lo = mid + 1; len = len - half - 1; } else len = half; }
  return lo;
}

/* StableSum (Neumaier), ref: include/NCrystal/internal/utils/NCMath.hh:526-537 */
typedef struct { double s, c; } orc_ssum;
static inline void orc_ssum_add(orc_ssum* t, double x)
{
  double n = t->s + x;
  t->c += (fabs(t->s) >= fabs(x)) ? ((t->s - n) + x) : ((x - n) + t->s);
  t->s = n;
}
static inline double orc_ssum_get(const orc_ssum* t) { return t->s + t->c; }

/* EnergyDomain::contains, ref: include/NCrystal/core/NCTypes.hh:435,833-842 */
static inline int orc_domain_contains(double lo, double hi, double e)
{
  int isnull = (lo > DBL_MAX) || (lo == hi);
  return !isnull && e >= lo && e <= hi;
}

/* pickRandIdxByWeight (n>=2), ref: src/utils/NCRandUtils.cc:198-220 */
static inline int orc_pick(double u, const double* cumul, int n)
{
  if (n < 5) {
    double choice = cumul[n - 1] * u;
    for (int i = 0; i < n; ++i) if (cumul[i] > choice) return i;
    return n - 1;
  }
  int i = orc_lower_bound(cumul, 0, n, cumul[n - 1] * u);
  return i < n - 1 ? i : n - 1;
}

/* ---- material model ---- */
typedef struct { int n; double threshold; const double* e2d; const double* fdm; } orc_pb;
typedef struct { int n; double msd[ORC_MAXEL], bixs[ORC_MAXEL]; } orc_elinc;
typedef struct { double sigma_free, ca, kT, mass_amu; } orc_fg;

typedef struct {            /* AlphaSampleInfo, ref: include/NCrystal/internal/sab/NCSABSamplerModels.hh:47-57 */
  double f_alpha, f_sval, f_logsval; int f_idx;
  double b_alpha, b_sval, b_logsval; int b_idx;
  double prob_front, prob_notback;
} orc_ainfo;

typedef struct {            /* SABSamplerAtE_Alg1 (or NoScatter when npts==0) */
  int npts, ibeta_off;
  double first_bin;         /* m_firstBinKinematicEndpointValue */
  double *x, *pdf, *cdf;    /* PointwiseDist m_betaSampler */
  orc_ainfo* infos;         /* npts-1 */
  double xs_check;          /* total xs recomputed by the integrator at this energy */
} orc_epoint;

typedef struct {
  double scale, kT, k_extension, k1, k2, egrid_margin, bound_xs;
  orc_fg ext;
  int negrid, nalpha, nbeta;
  const double *egrid, *xs, *alpha, *beta, *sab;
  double *logsab, *cumul;   /* CommonCache, ref: src/sab/NCSABIntegrator.cc:105-141 */
  orc_epoint* ep;
} orc_sab;

typedef struct { const double* data; int nm2; double a, invdelta; } orc_lut;
typedef struct {
  double threshold_ekin, cta, k1, k2, numint_accuracy;
  int nfam, nnormals;
  const double *fam_xsfact, *fam_inv2d, *fam_first, *normals;
  orc_lut sofcosd, evalcosx;
} orc_sc;

typedef struct { int kind; double scale, dom_lo, dom_hi; int idx; } orc_comp;

typedef struct {
  unsigned char* blob;      /* private copy */
  int ncomp, oriented;
  double dom_lo, dom_hi;
  orc_comp comp[ORC_MAXCOMP];
  orc_pb pb[ORC_MAXCOMP]; int npb;         /* (one slot per possible component of each kind) */
  orc_elinc elinc[ORC_MAXCOMP]; int nel;
  orc_fg fg[ORC_MAXCOMP]; int nfg;
  orc_sab sab[ORC_MAXCOMP]; int nsab;
  orc_sc sc; int nsc;
  char err[256];
} orc_material;

/* error flags raised where the reference throws */
enum { ORC_ERR_KIN = 1, ORC_ERR_OUTER = 2, ORC_ERR_INNER = 4, ORC_ERR_DISCARD = 8 };

/* oracle_iso.c */
double orc_xs_iso(const orc_material* M, double ekin, double* cumul, int* aux);
double orc_comp_xs_iso(const orc_material* M, int i, double ekin, int* aux);
void orc_comp_sample_iso(const orc_material* M, int i, int aux, double ekin, orc_rng* rng, double* eout, double* mu, int* err);
void orc_sample_iso(const orc_material* M, double ekin, orc_rng* rng, double* eout, double* mu, int* err);
int orc_sab_build(orc_sab* T);
void orc_sab_free(orc_sab* T);
/* oracle_sc.c */
typedef struct { double x, y, z; } orc_vec;
double orc_sc_xs(const orc_sc* S, double ekin, orc_vec dir, int* nentries);
void orc_sc_sample(const orc_sc* S, double ekin, orc_vec indir, int nentries, double total, orc_rng* rng, orc_vec* out);
orc_vec orc_rand_dir_given_mu(orc_rng* rng, double mu, orc_vec indir);
double orc_xs(const orc_material* M, double ekin, orc_vec dir, double* cumul, int* aux, double* sc_total);
void orc_sample(const orc_material* M, double ekin, orc_vec dir, orc_rng* rng, double* eout, orc_vec* out, int* err);

#endif
