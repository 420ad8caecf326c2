/* tests/capi_crng.c -- the reference's own pinned test of ncrystal_samplescatter_rs, restated: a tiny stateful
 * generator (31-bit LCG, one stream per struct) is handed to the library, which must consume it draw by draw like the
 * reference does -- every number it asks for is printed, so the output can be compared line by line with the
 * reference's log (tests/src/app_crng/main.c + test.log; the log is kept as tests/golden/ref_app_crng_test.log).
 * PowderBragg on a single-component material: exactly 3 numbers per scattering. */
#include "ncrystal_b200.h"
#include <stdint.h>
#include <stdio.h>

typedef struct { uint32_t data; } stream_t;

static double next_number( void* p )
{
  stream_t* s = (stream_t*)p;
  s->data = (uint32_t)( ( 1103515245u * s->data + 12345u ) % 2147483648u );
  const double v = s->data * ( 1.0 / 2147483648 );
  printf( "...custom_rng produces %g\n", v );
  return v;
}

static stream_t make_stream( uint32_t seed ) { stream_t s; s.data = 1789569706u + seed; return s; }

static void show( const stream_t* s1, const stream_t* s2 )
{
  printf( "stream1 state: %lu\n", (unsigned long)s1->data );
  if ( s2 ) printf( "stream2 state: %lu\n", (unsigned long)s2->data );
}

static void scatter( ncrystal_scatter_t sc, stream_t* s, int which )
{
  const double indir[3] = { 0.0, 0.0, 1.0 };
  double ekin_final, outdir[3];
  printf( "Requesting scatter (stream%d)\n", which );
  ncrystal_samplescatter_rs( next_number, s, sc, 0.025, (const double (*)[3])&indir, &ekin_final, (double (*)[3])&outdir );
}

int main( void )
{
  ncrystal_scatter_t sc = ncrystal_create_scatter( "stdlib::Al_sg225.ncmat;comp=bragg;dcutoff=2.2" );
  stream_t s1, s2;
  printf( "reset stream1 state\n" );
  s1 = make_stream( 12345 );
  show( &s1, 0 );
  scatter( sc, &s1, 1 ); show( &s1, 0 );
  scatter( sc, &s1, 1 ); show( &s1, 0 );
  printf( "reset stream1 state\n" );
  s1 = make_stream( 12345 );
  printf( "create stream2 as copy of stream1 state\n" );
  s2 = s1;
  show( &s1, &s2 );
  scatter( sc, &s1, 1 ); show( &s1, &s2 );
  scatter( sc, &s2, 2 ); show( &s1, &s2 );
  scatter( sc, &s1, 1 ); show( &s1, &s2 );
  ncrystal_unref( &sc );
  return 0;
}
