/* oracle/oracle_api.c -- TEST INFRASTRUCTURE (see oracle_common.h).  Exported entry points of
 * libncb200_oracle.so: load a compiled material, evaluate cross sections / sample scatterings one
 * neutron at a time with the replayed per-neutron Philox streams (philox_ref.h), optionally on
 * several host threads (for the "port" CPU baseline of bench.py). */
#include "oracle_common.h"
#include <pthread.h>
#include <stdio.h>

static const double* arr_after(const unsigned char* p, size_t hdr) { return (const double*)(p + hdr); }

void* orc_load(const void* blob, uint64_t nbytes)
{
  if (nbytes < sizeof(ncb_header_t)) return 0;
  orc_material* M = (orc_material*)calloc(1, sizeof(orc_material));
  M->blob = (unsigned char*)malloc(nbytes);
  memcpy(M->blob, blob, nbytes);
  const ncb_header_t* h = (const ncb_header_t*)M->blob;
  if (h->magic != NCB_MAGIC || h->version != NCB_VERSION || h->ncomp == 0 || h->ncomp > ORC_MAXCOMP) { free(M->blob); free(M); return 0; }
  M->ncomp = (int)h->ncomp; M->oriented = (int)h->oriented; M->dom_lo = h->dom_lo; M->dom_hi = h->dom_hi;
  for (int i = 0; i < M->ncomp; ++i) {
    const ncb_comp_t* c = &h->comp[i];
    orc_comp* k = &M->comp[i];
    k->kind = (int)c->kind; k->scale = c->scale; k->dom_lo = c->dom_lo; k->dom_hi = c->dom_hi;
    const unsigned char* p = M->blob + c->off;
    if (c->kind == NCB_KIND_POWDERBRAGG) {
      const ncb_powderbragg_t* q = (const ncb_powderbragg_t*)p;
      orc_pb* T = &M->pb[M->npb];
      T->n = (int)q->nplanes; T->threshold = q->threshold;
      T->e2d = arr_after(p, sizeof(*q)); T->fdm = T->e2d + q->nplanes;
      k->idx = M->npb++;
    } else if (c->kind == NCB_KIND_ELINC) {
      const ncb_elinc_t* q = (const ncb_elinc_t*)p;
      const double* a = arr_after(p, sizeof(*q));
      orc_elinc* T = &M->elinc[M->nel];
      T->n = (int)q->nelem;
      for (int j = 0; j < T->n; ++j) { T->msd[j] = a[j]; T->bixs[j] = a[q->nelem + j]; }
      k->idx = M->nel++;
    } else if (c->kind == NCB_KIND_FREEGAS) {
      const ncb_freegas_t* q = (const ncb_freegas_t*)p;
      orc_fg* T = &M->fg[M->nfg];
      T->sigma_free = q->sigma_free; T->ca = q->ca; T->kT = ORC_BOLTZMANN*q->temperature; T->mass_amu = q->mass_amu;
      k->idx = M->nfg++;
    } else if (c->kind == NCB_KIND_SAB) {
      const ncb_sab_t* q = (const ncb_sab_t*)p;
      const double* a = arr_after(p, sizeof(*q));
      orc_sab* T = &M->sab[M->nsab];
      T->scale = q->scale; T->kT = ORC_BOLTZMANN*q->temperature; T->k_extension = q->k_extension;
      T->k1 = q->k1; T->k2 = q->k2; T->egrid_margin = q->egrid_margin; T->bound_xs = q->bound_xs;
      T->ext.sigma_free = q->ext_sigma_free; T->ext.ca = q->ext_ca; T->ext.kT = ORC_BOLTZMANN*q->ext_temperature; T->ext.mass_amu = q->ext_mass_amu;
      T->negrid = (int)q->negrid; T->nalpha = (int)q->nalpha; T->nbeta = (int)q->nbeta;
      T->egrid = a; T->xs = a + q->negrid; T->alpha = T->xs + q->negrid; T->beta = T->alpha + q->nalpha; T->sab = T->beta + q->nbeta;
      if (orc_sab_build(T) != 0) { snprintf(M->err, sizeof(M->err), "SAB table build failed"); }
      k->idx = M->nsab++;
    } else if (c->kind == NCB_KIND_SCBRAGG) {
      const ncb_scbragg_t* q = (const ncb_scbragg_t*)p;
      const double* a = arr_after(p, sizeof(*q));
      orc_sc* S = &M->sc;
      S->threshold_ekin = q->threshold_ekin; S->cta = q->gos_cta; S->k1 = q->gos_circleint_k1; S->k2 = q->gos_circleint_k2;
      S->numint_accuracy = q->gos_numint_accuracy;
      S->nfam = (int)q->nfam; S->nnormals = (int)q->nnormals;
      S->fam_xsfact = a; S->fam_inv2d = a + q->nfam; S->fam_first = S->fam_inv2d + q->nfam; S->normals = S->fam_first + q->nfam + 1;
      S->sofcosd.data = S->normals + 3*q->nnormals; S->sofcosd.nm2 = (int)q->lut_sofcosd_n - 2; S->sofcosd.a = q->sofcosd_a; S->sofcosd.invdelta = q->sofcosd_invdelta;
      S->evalcosx.data = S->sofcosd.data + 2*q->lut_sofcosd_n; S->evalcosx.nm2 = (int)q->lut_evalcosx_n - 2; S->evalcosx.a = q->evalcosx_a; S->evalcosx.invdelta = q->evalcosx_invdelta;
      k->idx = 0; M->nsc = 1;
    }
  }
  return M;
}
void orc_free(void* vm)
{
  orc_material* M = (orc_material*)vm;
  if (!M) return;
  for (int i = 0; i < M->nsab; ++i) orc_sab_free(&M->sab[i]);
  free(M->blob); free(M);
}
const char* orc_error(void* vm) { return ((orc_material*)vm)->err; }
int orc_ncomp(void* vm) { return ((orc_material*)vm)->ncomp; }

/* total xs per SAB energy point recomputed by the integrator (compare with the reference's xs grid) */
int orc_sab_xscheck(void* vm, int comp, double* out)
{
  orc_material* M = (orc_material*)vm;
  if (comp < 0 || comp >= M->ncomp || M->comp[comp].kind != NCB_KIND_SAB) return -1;
  const orc_sab* T = &M->sab[M->comp[comp].idx];
  for (int i = 0; i < T->negrid; ++i) out[i] = T->ep[i].xs_check;
  return T->negrid;
}
/* same layout as refdrv_sab_sampler_dump */
int orc_sab_sampler_dump(void* vm, int comp, int iE, double* x, double* pdf, double* cdf, double* infos, double* meta)
{
  orc_material* M = (orc_material*)vm;
  if (comp < 0 || comp >= M->ncomp || M->comp[comp].kind != NCB_KIND_SAB) return -1;
  const orc_sab* T = &M->sab[M->comp[comp].idx];
  if (iE < 0 || iE >= T->negrid) return -1;
  const orc_epoint* ep = &T->ep[iE];
  for (int i = 0; i < ep->npts; ++i) { if (x) x[i] = ep->x[i]; if (pdf) pdf[i] = ep->pdf[i]; if (cdf) cdf[i] = ep->cdf[i]; }
  if (infos) for (int i = 0; i + 1 < ep->npts; ++i) {
    const orc_ainfo* f = &ep->infos[i]; double* o = infos + 10*i;
    o[0]=f->f_alpha; o[1]=f->f_sval; o[2]=f->f_logsval; o[3]=f->f_idx; o[4]=f->b_alpha; o[5]=f->b_sval; o[6]=f->b_logsval; o[7]=f->b_idx;
    o[8]=f->prob_front; o[9]=f->prob_notback;
  }
  if (meta) { meta[0] = ep->ibeta_off; meta[1] = ep->first_bin; }
  return ep->npts;
}

void orc_xs_iso_many(void* vm, const double* ekin, uint64_t n, double* out)
{
  const orc_material* M = (const orc_material*)vm;
  for (uint64_t i = 0; i < n; ++i) out[i] = orc_xs_iso(M, ekin[i], 0, 0);
}
void orc_sample_iso_many(void* vm, uint64_t seed, uint64_t first, const double* ekin, uint64_t n,
                         double* eout, double* mu, uint32_t* ndraws, int32_t* errs)
{
  const orc_material* M = (const orc_material*)vm;
  for (uint64_t i = 0; i < n; ++i) {
    orc_rng r; ncb_stream_init(&r, seed, first + i);
    int err = 0;
    orc_sample_iso(M, ekin[i], &r, &eout[i], &mu[i], &err);
    if (ndraws) ndraws[i] = r.ndraws;
    if (errs) errs[i] = err;
  }
}
void orc_xs_many(void* vm, const double* ekin, const double* ux, const double* uy, const double* uz, uint64_t n, double* out)
{
  const orc_material* M = (const orc_material*)vm;
  for (uint64_t i = 0; i < n; ++i) { orc_vec d = { ux[i], uy[i], uz[i] }; out[i] = orc_xs(M, ekin[i], d, 0, 0, 0); }
}
void orc_sample_many(void* vm, uint64_t seed, uint64_t first, const double* ekin, const double* ux, const double* uy, const double* uz,
                     uint64_t n, double* eout, double* ox, double* oy, double* oz, uint32_t* ndraws, int32_t* errs)
{
  const orc_material* M = (const orc_material*)vm;
  for (uint64_t i = 0; i < n; ++i) {
    orc_rng r; ncb_stream_init(&r, seed, first + i);
    int err = 0; orc_vec d = { ux[i], uy[i], uz[i] }, o;
    orc_sample(M, ekin[i], d, &r, &eout[i], &o, &err);
    ox[i] = o.x; oy[i] = o.y; oz[i] = o.z;
    if (ndraws) ndraws[i] = r.ndraws;
    if (errs) errs[i] = err;
  }
}

/* threaded CPU baseline ("port"): mode 0 = xs_iso, 1 = sample_iso; returns seconds of one pass */
typedef struct { void* vm; int mode; uint64_t seed; const double* ekin; uint64_t b, e; double *o0, *o1; } job;
static void* worker(void* p)
{
  job* j = (job*)p;
  if (j->mode == 0) orc_xs_iso_many(j->vm, j->ekin + j->b, j->e - j->b, j->o0 + j->b);
  else orc_sample_iso_many(j->vm, j->seed, j->b, j->ekin + j->b, j->e - j->b, j->o0 + j->b, j->o1 + j->b, 0, 0);
  return 0;
}
#include <time.h>
double orc_bench(void* vm, int mode, int nthreads, const double* ekin, uint64_t n, double* o0, double* o1)
{
  pthread_t th[256]; job jb[256];
  if (nthreads > 256) nthreads = 256;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) {
    jb[t].vm = vm; jb[t].mode = mode; jb[t].seed = 12345; jb[t].ekin = ekin; jb[t].o0 = o0; jb[t].o1 = o1;
    jb[t].b = n*(uint64_t)t/nthreads; jb[t].e = n*(uint64_t)(t+1)/nthreads;
    pthread_create(&th[t], 0, worker, &jb[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec);
}

/* 1/v absorption (AbsOOV::crossSectionIsotropic, ref: ncrystal_core/src/absoov/NCAbsOOV.cc:41-45) with the constant
 * the material compiler stored in the blob header; domain as AbsOOV::m_domain (:35-37). */
void orc_abs_xs_many(void* vm, const double* ekin, uint64_t n, double* out)
{
  const orc_material* M = (const orc_material*)vm;
  const double c = ((const ncb_header_t*)M->blob)->abs_c;
  for (uint64_t i = 0; i < n; ++i) {
    if (!(c > 0.0) || !(ekin[i] >= 0.0)) { out[i] = 0.0; continue; }   /* outside the (empty or [0,inf]) domain */
    const double s = sqrt(ekin[i]);
    out[i] = s ? c / s : INFINITY;
  }
}
