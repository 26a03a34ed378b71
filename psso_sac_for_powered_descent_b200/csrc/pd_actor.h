// pd_actor.h - shared-actor inference launchers (pd_actor.cu), used by pd_api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pd {

struct ActorArgs {
    const float *w1, *b1, *w2, *b2, *wm, *bm, *ws, *bs;   // device fp32, torch nn.Linear layouts
    const void *w2_img;                                   // bf16 shared-memory image of W2 (tensor-core path)
    int hidden;
    int deterministic;
    float max_action;
    unsigned int step;                                    // Philox counter (collection step index)
    unsigned long long seed;
};

size_t actor_tc_smem_bytes(int O, int A);
int actor_prep_w2(const float *w2, void *img, cudaStream_t st);
int actor_launch(int O, int A, const ActorArgs &p, const float *obs, float *act, float *mean_out, int B,
                 int use_tc, int n_sm, cudaStream_t st);

}  // namespace pd
