// Opt-in variant of the cooperative kernel for Bsim4 batches (config C4), selected at run time with S21_B4_FAST=1:
// the SAME kernel text as coop.cu, compiled once more with the evaluation's ~460 divisions written as a * rcp(b)
// (bsim4/bsim4_eval.hpp, S21_B4_RCPDIV). Measured on a B200, 21-stage ring x 100 points: 2048 instances 188 -> 141-151 ms,
// 256 instances 88.6 -> 70.4 ms (profiles/r02v_c4_rcpdiv.txt, r02w_c4_fast.txt); letting nvcc also contract a*b+c on
// top of it bought nothing (146 / 69.4 ms) and is not done. Results are within ~1e-10 of the default kernels on C4 instead
// of bit-identical to them, which is why this is not the default: tests/test_gpu.py::test_bsim4_fast_division_*.
#define S21_B4_RCPDIV 1
#define S21_COOP_FAST 1
#include "coop.cu"
