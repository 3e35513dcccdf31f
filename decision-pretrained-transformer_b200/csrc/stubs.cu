// Entry points declared in include/dpt_b200.h whose kernels are not built yet in this round:
// they fail loudly with DPT_ERR_UNSUPPORTED (never a silent fallback).  This file shrinks as the
// kernels land and disappears when the header is fully implemented.
#include "common.cuh"

#define DPT_STUB(name)                                                  \
  dpt::set_error(name ": not implemented in this build of libdpt_b200"); \
  return DPT_ERR_UNSUPPORTED;

extern "C" int dpt_gpt2_create(const dpt_gpt2_weights_t*, dpt_gpt2_t**, void*) { DPT_STUB("dpt_gpt2_create") }
extern "C" int dpt_gpt2_destroy(dpt_gpt2_t*) { DPT_STUB("dpt_gpt2_destroy") }
extern "C" int dpt_gpt2_forward(dpt_gpt2_t*, const float*, const float*, const float*, const float*, const float*, int,
                                int, int, int, int, float*, void*) {
  DPT_STUB("dpt_gpt2_forward")
}
extern "C" uint64_t dpt_gpt2_online_kv_bytes(const dpt_gpt2_t*, int, int, int) { return 0; }
extern "C" int dpt_gpt2_online_loop(dpt_gpt2_t*, const float*, double, int, uint64_t, uint64_t, int, int, int, void*,
                                    uint64_t, float*, float*, float*, float*, float*, double*,
                                    const dpt_gpt2_online_inject_t*, const dpt_gpt2_online_dump_t*, void*) {
  DPT_STUB("dpt_gpt2_online_loop")
}
