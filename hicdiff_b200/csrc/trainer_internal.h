// Shared between trainer.cu (hd_trainer C ABI + the hicedrn_Diff step) and unet_trainer.cu (the Unet step).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/hicdiff_b200.h"
#include "kernels.h"

namespace hd {
int set_error(const char* msg);   // plan.cu: fills the thread-local message behind hd_last_error()

inline int tfail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return set_error(buf);
}

struct TParam {
    const float* w = nullptr;
    float* g = nullptr;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

struct TOp {
    std::function<cudaError_t(cudaStream_t)> fn;
    std::string tag;
    const char* kernel = "";
    double flops = 0;
    // graph capture only (the eager path runs everything on the caller's stream, in order): side = 1 launches the op on the
    // trainer's side stream, forked from the main stream at this point of the list (so it sees everything issued before it);
    // join = 1 makes a main-stream op wait for all side work issued so far.  Lets HBM-bound reductions overlap the tensor-bound GEMMs.
    int side = 0;
    int join = 0;
    // >= 0: a gradient-bucket marker (no kernel): every gradient of bucket `bucket` is final once the ops before it -- side
    // stream included -- have run; the step records the bucket's event here so a communication stream can start its all-reduce
    int bucket = -1;
};
}  // namespace hd

struct hd_trainer {
    hd_config cfg;
    int device = 0, num_sms = 148, B = 0, nb = 0;
    bool finalized = false;
    std::map<std::string, hd::TParam> p;
    std::vector<void*> allocs;
    size_t bytes = 0;
    std::vector<hd::TOp> ops;
    float *x = nullptr, *cond = nullptr, *time = nullptr, *target = nullptr, *weight = nullptr, *eps = nullptr, *d_eps = nullptr,
          *loss = nullptr;
    int loss_kind = 1;
    // the whole step as a CUDA graph per loss kind: step 1 runs eagerly (validates every launch), step 2 is captured on a
    // trainer-private stream, later steps replay (HD_TRAIN_GRAPH=0 keeps the eager path)
    int eager_steps[2] = {0, 0};
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    cudaStream_t cap_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // data-parallel training (hd_trainer_set_grad_buckets): bucket k is complete when the backward reaches the first module whose
    // name starts with bucket_prefix[k]; bucket_ev[k] is recorded there (an external event-record node inside the step's graph)
    std::vector<std::string> bucket_prefix;
    std::vector<cudaEvent_t> bucket_ev;
};

namespace hd {

template <typename T>
int dalloc(hd_trainer* t, T** out, size_t bytes, bool zero = false) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
    if (e != cudaSuccess) return tfail("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    if (zero) cudaMemset(q, 0, bytes ? bytes : 16);
    t->allocs.push_back(q);
    t->bytes += bytes;
    *out = static_cast<T*>(q);
    return 0;
}

inline const TParam* find_p(const hd_trainer* t, const std::string& k, std::initializer_list<int64_t> shape) {
    auto it = t->p.find(k);
    if (it == t->p.end()) { tfail("parameter '%s' was not bound", k.c_str()); return nullptr; }
    if (it->second.shape != std::vector<int64_t>(shape)) { tfail("parameter '%s' has an unexpected shape", k.c_str()); return nullptr; }
    return &it->second;
}

int build_unet_trainer(hd_trainer* t);   // unet_trainer.cu

}  // namespace hd
