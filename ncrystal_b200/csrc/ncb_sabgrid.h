// ncb_sabgrid.h -- energy grid of an S(alpha,beta) leaf determined from the scattering kernel alone.
// Restates SABIntegrator::Impl::setupEnergyGrid / determineEMin / determineEMax (ref: src/sab/NCSABIntegrator.cc:
// 147-283) and geomspace (src/utils/NCMath.cc:44-66).  Host code; `sigmaAt(energies)` returns the integrated cross
// section of the kernel at each energy -- in the product it runs the table-build kernels over the whole probe
// sequence at once (ncb_lib.cu), in the CPU test build it calls the same functions on the host (tests/hostsim).
//
// The reference probes sigma(E) one energy at a time: determineEMax walks E *= 0.95 down from the kinematic limit of
// the table until the distance to the free-gas cross section grows again (<= 180 points), determineEMin halves E
// until sqrt(E)*sigma(E) is flat to 1e-3 (<= 100 points).  Both probe sequences are fixed in advance, so all points
// are integrated in one batch and the reference's stopping rule is then applied to the results.
#pragma once
#include "ncb_phys_basic.cuh"
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

namespace ncb {

  template <class SigmaAt, class Warn>
  // req_emin / req_emax: the caller's request in the {emin, emax, npts} form of the reference (an NCMAT file's
  // "egrid" line), 0 = determine automatically.
  inline std::vector<double> sabDetermineEnergyGrid( int npts, double kT, double beta_min, double alpha_max, double suggested_emax,
                                                     double req_emin, double req_emax,
                                                     const FreeGasT& ext, SigmaAt&& sigmaAt, Warn&& warn )
  {
    if ( !( req_emin >= 0.0 ) || !( req_emax >= 0.0 ) || !( req_emax == 0.0 || req_emin == 0.0 || req_emax > req_emin ) )
      throw std::runtime_error( "SABIntegrator invalid energy grid. Values for emin/emax must fullfil 0<emin<emax or be 0 indicating automatic determination." );
    double emax = req_emax;
    if ( emax == 0.0 && suggested_emax > 0.0 ) {
      if ( req_emin != 0.0 && suggested_emax <= req_emin )
        throw std::runtime_error( "SABIntegrator invalid energy grid: emin must be below the table's suggested Emax" );
      emax = suggested_emax;
    }
    if ( emax == 0.0 ) {
      // the kinematic curve through (alpha_max, beta_min) bounds the table's energy range
      const double emax_upper_limit = kT*( beta_min-alpha_max )*( beta_min-alpha_max )/( 4*alpha_max );
      std::vector<double> probe;
      const double elow = emax_upper_limit*1e-4;
      for ( double e = emax_upper_limit; e > elow; e *= 0.95 ) probe.push_back( e );
      const std::vector<double> xs = sigmaAt( probe );
      double prev = kInf;
      for ( size_t k = 0; k < probe.size(); ++k ) {
        const double dist = std::fabs( xs[k] - fgXS( ext, probe[k] ) );
        if ( dist > prev ) { emax = 0.95*probe[k]; break; }
        prev = dist;
      }
      if ( !( emax > 0.0 ) ) {
        emax = 0.5*emax_upper_limit;
        warn( "Algorithm searching for suitable Emax value at which to end SAB energy grid failed to provide reasonable result. Using crude guess." );
      }
    }
    double emin = req_emin > 0.0 ? req_emin : -1.0;
    if ( emin < 0.0 ) {
      const double e_upp = std::min( emax*0.01, 0.01*kT );
      std::vector<double> probe;
      probe.push_back( e_upp*0.9 );
      while ( 0.5*probe.back() > 1e-30*e_upp ) probe.push_back( 0.5*probe.back() );
      const std::vector<double> xs = sigmaAt( probe );
      double e_old = probe[0], f_old = std::sqrt( probe[0] )*xs[0];
      for ( size_t k = 1; k < probe.size(); ++k ) {
        const double f_new = std::sqrt( probe[k] )*xs[k];
        if ( f_new == 0.0 ) {
          warn( "Encountered sqrt(E)*sigma(E)=0 while searching for suitable Emin value at which to start SAB energy grid. Will revert to using Emin=0.001*Emax." );
          emin = 0.001*e_upp;
          break;
        }
        if ( std::fabs( f_old/f_new - 1.0 ) < 1e-3 ) { emin = e_old; break; }
        e_old = probe[k]; f_old = f_new;
      }
      if ( emin < 0.0 ) emin = std::min( e_old, e_upp*0.01 );
    }
    if ( !( emin > 0.0 ) || !( emax > emin ) || npts < 10 )
      throw std::runtime_error( "SABIntegrator invalid energy grid" );
    std::vector<double> egrid( npts );
    double start = std::log10( emin );
    const double stop = std::log10( emax ), interval = ( stop - start )/( npts - 1 );
    for ( int k = 0; k < npts; ++k ) { egrid[k] = std::pow( 10.0, start ); start += interval; }
    egrid[npts-1] = std::pow( 10.0, stop );
    egrid[0] = emin; egrid[npts-1] = emax;
    for ( int k = 1; k < npts; ++k )
      if ( !( egrid[k] > egrid[k-1] ) )
        throw std::runtime_error( "SABIntegrator invalid energy grid - must be sorted with non-repeated and positive values." );
    return egrid;
  }

}
