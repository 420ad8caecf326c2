/* tests/capi_caller.c -- a plain C caller written against include/ncrystal_b200.h only, with the call sequence of
 * the reference's examples/ncrystal_example_c.c (polycrystal Al, then the Ge single crystal): what a C user of
 * NCrystal's C-API compiles and links unchanged against libncrystal_b200.so.  Prints "key value" lines for the test. */
#include "ncrystal_b200.h"
#include <stdio.h>

int main(void)
{
  ncrystal_scatter_t pc, sc;
  ncrystal_process_t pc_proc, sc_proc;
  double wl, ekin, xsect, ekin_final, cos_scat_angle;
  double outdir[3];
  unsigned i, nelastic = 0;
  const double dir1[3] = { 0., 1., 1. };
  const double dir2[3] = { 1., 1., 0. };
  double e_many[4], xs_many[4], eo_many[8], mu_many[8];

  pc = ncrystal_create_scatter_builtinrng( "Al_sg225.ncmat;temp=293.15K", 2026 );
  if ( !pc.internal || ncrystal_error() ) { printf( "error %s\n", ncrystal_lasterror() ); return 1; }
  pc_proc = ncrystal_cast_scat2proc( pc );
  wl = 2.5;
  ekin = ncrystal_wl2ekin( wl );
  ncrystal_crosssection_nonoriented( pc_proc, ekin, &xsect );
  printf( "al_xs_2.5Aa %.17g\n", xsect );
  for ( i = 0; i < 20; ++i ) {
    ncrystal_samplescatterisotropic( pc, ekin, &ekin_final, &cos_scat_angle );
    if ( ekin_final == ekin ) ++nelastic;
    if ( !( cos_scat_angle >= -1.0 && cos_scat_angle <= 1.0 ) || !( ekin_final >= 0.0 ) ) { printf( "bad sample\n" ); return 1; }
  }
  printf( "al_nelastic_of_20 %u\n", nelastic );
  for ( i = 0; i < 4; ++i ) e_many[i] = ncrystal_wl2ekin( 1.0 + i );
  ncrystal_crosssection_nonoriented_many( pc_proc, e_many, 4, 1, xs_many );
  ncrystal_samplescatterisotropic_many( pc, e_many, 4, 2, eo_many, mu_many );
  printf( "al_xs_many %.17g %.17g %.17g %.17g\n", xs_many[0], xs_many[1], xs_many[2], xs_many[3] );
  printf( "al_refcount %d\n", ncrystal_refcount( &pc ) );

  wl = 1.540;
  ekin = ncrystal_wl2ekin( wl );
  sc = ncrystal_create_scatter( "Ge_sg227.ncmat;mos=40arcsec;dir1=@crys_hkl:5,1,1@lab:0,0,1;dir2=@crys_hkl:0,-1,1@lab:0,1,0" );
  if ( !sc.internal || ncrystal_error() ) { printf( "error %s\n", ncrystal_lasterror() ); return 1; }
  sc_proc = ncrystal_cast_scat2proc( sc );
  ncrystal_crosssection( sc_proc, ekin, &dir1, &xsect );
  printf( "ge_xs_dir1 %.17g\n", xsect );
  ncrystal_crosssection( sc_proc, ekin, &dir2, &xsect );
  printf( "ge_xs_dir2 %.17g\n", xsect );
  ncrystal_samplescatter( sc, ekin, &dir1, &ekin_final, &outdir );
  printf( "ge_outdir_norm2 %.17g\n", outdir[0]*outdir[0] + outdir[1]*outdir[1] + outdir[2]*outdir[2] );
  printf( "ge_isnonoriented %d\n", ncrystal_isnonoriented( sc_proc ) );

  ncrystal_unref( &pc );
  ncrystal_unref( &sc );
  printf( "handles_cleared %d\n", ( pc.internal == 0 && sc.internal == 0 ) ? 1 : 0 );
  return ncrystal_error() ? 1 : 0;
}
