// ipm_tiny.cu — second instantiation of the single-CTA interior-point kernel (ipm_small.cu) for frontier batches of relaxations
// whose SDP blocks all have order <= 16 (example_small, example_TT, example_MkP): CTAs of 256 threads with 41 KB of shared memory,
// so that four nodes share an SM.  Only the constants differ; see the SDPK_VARIANT_TINY block at the top of ipm_small.cu.
#define SDPK_VARIANT_TINY 1
#include "ipm_small.cu"
