// ncb_replay.h -- interface between ncb_lib.cu and ncb_replay.cu (one neutron sampled with the caller's numbers).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
constexpr int kReplayMaxNumbers = 512;   // numbers of the caller's generator one scattering may consume
void ncb_replay_sample_one( const void* material, size_t material_bytes, int device, const double* u, uint32_t nu,
                            double ekin, const double* dir, double* out4, uint32_t* ndraws, int* overrun, int* err,
                            void* stream );
