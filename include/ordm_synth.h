/*
 * ordm_synth.h -- counter-based synthetic input generator (host + device, bit-identical).
 *
 * OpenRDM ships only the two H2/STO-3G fixtures (tests/h2_t{svwn,pbe}_sto3g);
 * BASELINE.json's configs 3-5 are synthetic grids of named shapes.  Every value is a pure
 * function of (seed, stream, global index): integer hashing followed by exactly-rounded
 * IEEE-754 operations only (no libm), so the host, the CUDA fill kernels and any grid shard
 * regenerate the *same bits* for the same global grid point.  That is what lets the CPU
 * oracle be run on a subsample of a 5M / 20M-point grid and compared point by point.
 *
 *   phi  (p,i) = env(p) * (u01(seed, ORDM_STREAM_PHI,   i*npts_total + p) - 0.5)
 *   phi_c(p,i) = env(p) * (u01(seed, ORDM_STREAM_PHI+c, i*npts_total + p) - 0.5)   c=1,2,3 (x,y,z)
 *   w    (p)   = u01(seed, ORDM_STREAM_W, p) * (1.0 / npts_total)
 *   env  (p)   = 1 for ~90 % of the points, 2^-k (k = 0..46 uniform) for the rest: drives
 *                rho through the 1e-20 thresholds of mcpdft.cc:319-321 / functional.cc:74-93.
 */
#ifndef ORDM_SYNTH_H
#define ORDM_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ORDM_HD __host__ __device__ __forceinline__
#else
#define ORDM_HD static inline
#endif

enum {
  ORDM_STREAM_PHI = 1, /* +0 phi, +1 phi_x, +2 phi_y, +3 phi_z */
  ORDM_STREAM_ENV = 5,
  ORDM_STREAM_W = 6,
  ORDM_STREAM_RDM = 16 /* host-side RDM generator streams start here */
};

/* SplitMix64 finaliser (Steele, Lea & Flood 2014). */
ORDM_HD uint64_t ordm_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

ORDM_HD uint64_t ordm_stream_key(uint64_t seed, uint32_t stream) {
  return ordm_mix64(seed * 0xD1342543DE82EF95ULL + (uint64_t)stream * 0x2545F4914F6CDD1DULL);
}

/* uniform in [0,1) with 53 random bits */
ORDM_HD double ordm_u01_key(uint64_t key, uint64_t idx) {
  uint64_t r = ordm_mix64(key ^ (idx * 0x9E3779B97F4A7C15ULL));
  return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

ORDM_HD double ordm_u01(uint64_t seed, uint32_t stream, uint64_t idx) {
  return ordm_u01_key(ordm_stream_key(seed, stream), idx);
}

/* amplitude envelope of grid point p (global index) */
ORDM_HD double ordm_env_key(uint64_t env_key, uint64_t p) {
  double r = ordm_u01_key(env_key, p);
  if (r < 0.9) return 1.0;
  int k = (int)((r - 0.9) * 470.0); /* 0..46 */
  if (k > 46) k = 46;
  if (k < 0) k = 0;
  union {
    uint64_t u;
    double d;
  } b;
  b.u = (uint64_t)(1023 - k) << 52; /* exactly 2^-k */
  return b.d;
}

ORDM_HD double ordm_synth_phi(uint64_t comp_key, double env, uint64_t npts_total, uint64_t p, uint64_t i) {
  return env * (ordm_u01_key(comp_key, i * npts_total + p) - 0.5);
}

ORDM_HD double ordm_synth_w(uint64_t w_key, double inv_npts_total, uint64_t p) {
  return ordm_u01_key(w_key, p) * inv_npts_total;
}

#endif /* ORDM_SYNTH_H */
