/* oracle/philox_ref.h -- TEST INFRASTRUCTURE (oracle side).  CPU statement of the
 * per-neutron counter-based random stream the product uses on the device, so the
 * reference can be replayed with *identical* uniforms through its own RNG hook
 * (NCrystal::RNGStream subclass, ref: ncrystal_core/include/NCrystal/interfaces/NCRNG.hh;
 * C-API hook ncrystal_samplescatter_rs, ncrystal.h:792).
 *
 * Generator: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers:
 * as easy as 1, 2, 3", SC'11) -- restated from the published algorithm.
 *
 * Stream definition (must match ncrystal_b200/csrc/ncb_rng.cuh):
 *   key      = (seed_lo, seed_hi)
 *   counter  = (index_lo, index_hi, k>>1, sid)   index = global neutron index, sid = stream id
 *                                                (0 for a fresh handle, k for its k'th clone)
 *   draw k   = 64-bit word (k&1) of that block: w0 = r1:r0, w1 = r3:r2
 *   uniform  = randUInt64ToFP01(word)  in (0,1]  (ref: NCDefs.hh:1308-1330)
 */
#ifndef NCB_PHILOX_REF_H
#define NCB_PHILOX_REF_H
#include <stdint.h>

static inline void ncb_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* uint64 -> (0,1], ref: NCDefs.hh:1308-1330 (randUInt64ToFP01) */
static inline double ncb_u64_to_fp01(uint64_t x)
{
  const double r1 = (double)(x >> 11) * 0x1.0p-53;
  const double r2 = (double)(x & 0x7FF) * 0x1.0p-64;
  return (1.0 - r1) - r2;
}

typedef struct {
  uint32_t key[2];
  uint32_t ctr[4];   /* ctr[2] = block number */
  uint32_t buf[4];
  uint32_t ndraws;   /* draws consumed so far */
} ncb_stream_t;

static inline void ncb_stream_init_sid(ncb_stream_t* s, uint64_t seed, uint64_t index, uint32_t sid)
{
  s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
  s->ctr[0] = (uint32_t)index; s->ctr[1] = (uint32_t)(index >> 32);
  s->ctr[2] = 0; s->ctr[3] = sid;
  s->ndraws = 0;
}

static inline void ncb_stream_init(ncb_stream_t* s, uint64_t seed, uint64_t index)
{
  ncb_stream_init_sid(s, seed, index, 0);
}

static inline double ncb_stream_next(ncb_stream_t* s)
{
  const uint32_t k = s->ndraws++;
  uint64_t w;
  if ((k & 1u) == 0) {
    s->ctr[2] = k >> 1;
    ncb_philox4x32_10(s->ctr, s->key, s->buf);
    w = ((uint64_t)s->buf[1] << 32) | s->buf[0];
  } else {
    w = ((uint64_t)s->buf[3] << 32) | s->buf[2];
  }
  return ncb_u64_to_fp01(w);
}

#endif
