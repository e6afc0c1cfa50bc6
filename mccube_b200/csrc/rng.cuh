// Step-key derivation shared by the index / weighted-sampling kernels.
#pragma once
#include "common.cuh"

namespace mccb {

constexpr int MAX_ROUNDS = 8;

// Derived keys, produced on the device so key / t may live in device memory.
struct DerivedKeys {
  u32x2 step_key;               // key handed to jr.choice
  u32x2 round_sub[MAX_ROUNDS];  // sort-key generator of each _shuffle round
  u32x2 randint_hi, randint_lo; // k1, k2 of randint
};

// enqueue the 1-thread derivation kernel
void launch_derive_keys(const mccb200_rng& rng, int rounds, DerivedKeys* out, cudaStream_t st);

}  // namespace mccb
