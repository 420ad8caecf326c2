/* tests/capi_multigpu.c -- a plain C99 caller (the shape of an OpenMC / McStas integration: one process, host arrays)
 * that uses several GPUs through the C ABI alone: ncb200_set_devices(n), then the reference's own *_many entry points.
 * Checks that the n-device results are bit-identical to the one-device results and that the NCCL-merged tally
 * histogram equals the one-device histogram.   usage: capi_multigpu <cfg> <ndev> [n]   (exit 0 = pass) */
#include "ncrystal_b200.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime( CLOCK_MONOTONIC, &t ); return t.tv_sec + 1e-9*t.tv_nsec; }

int main( int argc, char** argv )
{
  const char* cfg = argc > 1 ? argv[1] : "Al_sg225.ncmat;temp=293.15K";
  const int ndev = argc > 2 ? atoi( argv[2] ) : 0;
  const unsigned long n = argc > 3 ? strtoul( argv[3], 0, 10 ) : 8000003ul;
  const unsigned nbins = 200;
  double *e, *xs1, *xsn, *eo1, *eon, *mu1, *mun, *h1, *hn;
  unsigned long i;
  unsigned long long seed = 88172645463325252ull;
  int got, bad = 0;
  double t0, t1, tn;
  ncrystal_scatter_t sc;
  ncrystal_process_t pr;

  ncrystal_sethaltonerror( 0 );
  e = malloc( n*sizeof(double) ); xs1 = malloc( n*sizeof(double) ); xsn = malloc( n*sizeof(double) );
  eo1 = malloc( n*sizeof(double) ); eon = malloc( n*sizeof(double) ); mu1 = malloc( n*sizeof(double) ); mun = malloc( n*sizeof(double) );
  h1 = calloc( nbins + 2, sizeof(double) ); hn = calloc( nbins + 2, sizeof(double) );
  for ( i = 0; i < n; ++i ) {          /* log-uniform 1e-5 .. 10 eV (xorshift64) */
    seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
    e[i] = pow( 10.0, -5.0 + 6.0*( (double)( seed >> 11 ) / 9007199254740992.0 ) );
  }
  sc = ncrystal_create_scatter_builtinrng( cfg, 2024 );
  if ( !sc.internal || ncrystal_error() ) { printf( "create failed: %s\n", ncrystal_lasterror() ); return 2; }
  pr = ncrystal_cast_scat2proc( sc );

  /* one device */
  ncb200_set_devices( 1 );
  ncb200_set_rng_stream( sc, 2024, 0, 0 );
  ncrystal_crosssection_nonoriented_many( pr, e, n, 1, xs1 );
  ncrystal_samplescatterisotropic_many( sc, e, n, 1, eo1, mu1 );        /* warm-up + reference results */
  ncb200_set_rng_stream( sc, 2024, 0, 0 );
  t0 = now();
  ncrystal_crosssection_nonoriented_many( pr, e, n, 1, xs1 );
  ncrystal_samplescatterisotropic_many( sc, e, n, 1, eo1, mu1 );
  t1 = now() - t0;
  ncb200_tally_hist_many( mu1, 0, n, -1.0, 1.0, nbins, h1, 0 );

  /* n devices */
  got = ncb200_set_devices( ndev );
  if ( got < 0 || ncrystal_error() ) { printf( "set_devices failed: %s\n", ncrystal_lasterror() ); return 2; }
  ncb200_set_rng_stream( sc, 2024, 0, 0 );
  ncrystal_crosssection_nonoriented_many( pr, e, n, 1, xsn );
  ncrystal_samplescatterisotropic_many( sc, e, n, 1, eon, mun );        /* (first call uploads the tables to the other devices) */
  ncb200_set_rng_stream( sc, 2024, 0, 0 );
  t0 = now();
  ncrystal_crosssection_nonoriented_many( pr, e, n, 1, xsn );
  ncrystal_samplescatterisotropic_many( sc, e, n, 1, eon, mun );
  tn = now() - t0;
  ncb200_tally_hist_many( mun, 0, n, -1.0, 1.0, nbins, hn, 0 );
  if ( ncrystal_error() ) { printf( "error: %s\n", ncrystal_lasterror() ); return 2; }

  bad += memcmp( xs1, xsn, n*sizeof(double) ) != 0;
  bad += memcmp( eo1, eon, n*sizeof(double) ) != 0;
  bad += memcmp( mu1, mun, n*sizeof(double) ) != 0;
  bad += memcmp( h1, hn, ( nbins + 2 )*sizeof(double) ) != 0;
  { double tot = 0; for ( i = 0; i < nbins + 2; ++i ) tot += hn[i]; bad += ( tot != (double)n ); }
  printf( "{\"cfg\": \"%s\", \"n\": %lu, \"devices\": %d, \"identical_to_one_device\": %s, \"neutrons_per_s_1dev\": %.4g, "
          "\"neutrons_per_s_ndev\": %.4g}\n", cfg, n, got, bad ? "false" : "true", n/t1, n/tn );
  ncrystal_unref( &sc );
  return bad ? 1 : 0;
}
