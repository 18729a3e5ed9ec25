// Production precision (fp32).  FMA contraction allowed.
#include "gsm_kernels_lane.cuh"
#include "gsm_kernels_team.cuh"
#include "gsm_kernels_wide.cuh"
#define GSM_REAL float
#define GSM_SFX(name) name##_f32
#include "gsm_launch.inl"
