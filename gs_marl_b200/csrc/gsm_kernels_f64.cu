// Verification precision (fp64).  Compile this file with -fmad=false: every operation
// rounds once, in the order SPEC.md writes it, like the CPU oracle.
#include "gsm_kernels_lane.cuh"
#include "gsm_kernels_team.cuh"
#include "gsm_kernels_wide.cuh"
#define GSM_REAL double
#define GSM_SFX(name) name##_f64
#include "gsm_launch.inl"
