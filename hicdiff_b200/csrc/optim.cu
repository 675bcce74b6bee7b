// hd_adam_*: the optimiser of the training loop (train.py:111 `torch.optim.Adam(diffusion.parameters(), lr=2e-5)`, stepped at
// train.py:129) as ONE launch over all parameter tensors.  State (exp_avg, exp_avg_sq) lives in two flat fp32 buffers in
// parameter order; the chunk table maps each block to <= ADAM_CHUNK elements of one tensor.
#include <cstring>
#include <vector>

#include "trainer_internal.h"

using namespace hd;

struct hd_adam {
    std::vector<float*> params;
    std::vector<int64_t> numels;
    std::vector<const float*> grads;      // the gradient pointers the device table was built for
    std::vector<AdamChunk> host;          // chunk table, host copy
    AdamChunk* staging = nullptr;         // pinned
    AdamChunk* table = nullptr;           // device
    float* exp_avg = nullptr;
    float* exp_avg_sq = nullptr;
    int64_t total = 0;
    int64_t step = 0;
    int device = 0;                       // the device the state lives on (current at hd_adam_create)
    bool zeroed = false;                  // the moments are zero-filled by the first step, on ITS stream
};

namespace {
// Run the body with the optimiser's device current (the caller may have another one selected), restore on exit.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};
}  // namespace

extern "C" {

int hd_adam_create(const void* const* params, const int64_t* numels, int32_t nparams, hd_adam** out) {
    if (!out) return tfail("hd_adam_create: null out");
    *out = nullptr;
    if (nparams < 0 || (nparams > 0 && (!params || !numels))) return tfail("hd_adam_create: bad argument");
    hd_adam* a = new hd_adam();
    if (nparams > 0) {
        // the state lives next to the parameters: take the device from the first parameter pointer
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, params[0]) == cudaSuccess && at.type == cudaMemoryTypeDevice) a->device = at.device;
        else cudaGetDevice(&a->device);
    } else {
        cudaGetDevice(&a->device);
    }
    DeviceGuard guard(a->device);
    for (int i = 0; i < nparams; ++i) {
        if (numels[i] < 0 || (numels[i] > 0 && !params[i])) {
            delete a;
            return tfail("hd_adam_create: parameter %d is null or has a negative size", i);
        }
        a->params.push_back(static_cast<float*>(const_cast<void*>(params[i])));
        a->numels.push_back(numels[i]);
        // keep every tensor's state 16-byte aligned so the vector path applies whenever the tensor itself is aligned
        const int64_t off = a->total;
        for (int64_t c = 0; c < numels[i]; c += ADAM_CHUNK) {
            AdamChunk ch;
            ch.param = a->params.back() + c;
            ch.grad = nullptr;
            ch.state_off = off + c;
            ch.n = static_cast<int>(numels[i] - c < ADAM_CHUNK ? numels[i] - c : ADAM_CHUNK);
            ch.pad = i;                    // parameter index: which gradient pointer this chunk reads
            a->host.push_back(ch);
        }
        a->total = off + ((numels[i] + 3) / 4) * 4;
    }
    a->grads.assign(nparams, nullptr);
    const size_t tb = a->host.size() * sizeof(AdamChunk), sb = static_cast<size_t>(a->total) * sizeof(float);
    cudaError_t e = cudaSuccess;
    if (tb) {
        e = cudaMallocHost(&a->staging, tb);
        if (e == cudaSuccess) e = cudaMalloc(&a->table, tb);
    }
    if (e == cudaSuccess && sb) e = cudaMalloc(&a->exp_avg, sb);
    if (e == cudaSuccess && sb) e = cudaMalloc(&a->exp_avg_sq, sb);
    if (e != cudaSuccess) {
        hd_adam_destroy(a);
        return tfail("hd_adam_create: %s", cudaGetErrorString(e));
    }
    *out = a;
    return 0;
}

int hd_adam_step(hd_adam* a, const void* const* grads, double lr, double beta1, double beta2, double eps, double weight_decay,
                 void* stream) {
    NvtxRange range("hd_adam_step");
    if (!a) return tfail("hd_adam_step: null optimiser");
    if (!grads && !a->params.empty()) return tfail("hd_adam_step: null gradient list");
    if (!(lr >= 0) || !(beta1 >= 0 && beta1 < 1) || !(beta2 >= 0 && beta2 < 1) || !(eps >= 0) || !(weight_decay >= 0))
        return tfail("hd_adam_step: invalid hyper-parameter (lr %g, betas %g %g, eps %g, weight_decay %g)", lr, beta1, beta2, eps,
                     weight_decay);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DeviceGuard guard(a->device);
    if (!a->zeroed && a->total > 0) {
        // ordered on the caller's stream (a legacy-stream cudaMemset is not ordered against a non-blocking stream)
        const size_t sb = static_cast<size_t>(a->total) * sizeof(float);
        cudaError_t ez = cudaMemsetAsync(a->exp_avg, 0, sb, s);
        if (ez == cudaSuccess) ez = cudaMemsetAsync(a->exp_avg_sq, 0, sb, s);
        if (ez != cudaSuccess) return tfail("hd_adam_step: %s", cudaGetErrorString(ez));
    }
    a->zeroed = true;
    const size_t np = a->params.size();
    bool changed = false;
    for (size_t i = 0; i < np; ++i) {
        if (!grads[i] && a->numels[i] > 0) return tfail("hd_adam_step: gradient %zu is null (every parameter must have one)", i);
        if (grads[i] != a->grads[i]) changed = true;
    }
    if (changed && !a->host.empty()) {
        // the previous step may still be reading the table / the staging copy: drain the stream before rewriting them
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return tfail("hd_adam_step: %s", cudaGetErrorString(e));
        for (size_t i = 0; i < np; ++i) a->grads[i] = static_cast<const float*>(grads[i]);
        int64_t first = 0;                 // element offset of the chunk inside its tensor
        int prev = -1;
        for (AdamChunk& ch : a->host) {
            if (ch.pad != prev) { prev = ch.pad; first = 0; }
            ch.grad = a->grads[ch.pad] + first;
            first += ch.n;
        }
        std::memcpy(a->staging, a->host.data(), a->host.size() * sizeof(AdamChunk));
        e = cudaMemcpyAsync(a->table, a->staging, a->host.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return tfail("hd_adam_step: %s", cudaGetErrorString(e));
    }
    a->step += 1;
    cudaError_t e = adam_step_run(a->table, static_cast<int>(a->host.size()), a->exp_avg, a->exp_avg_sq, lr, beta1, beta2, eps,
                                  weight_decay, a->step, s);
    if (e != cudaSuccess) {
        a->step -= 1;
        return tfail("hd_adam_step: %s", cudaGetErrorString(e));
    }
    return 0;
}

int hd_adam_state(hd_adam* a, int32_t index, float** exp_avg, float** exp_avg_sq, int64_t* step) {
    if (!a) return tfail("hd_adam_state: null optimiser");
    if (index < 0 || index >= static_cast<int32_t>(a->params.size())) return tfail("hd_adam_state: index %d out of range", index);
    if (!a->zeroed && a->total > 0) {
        // state requested before the first step (checkpoint restore): define it now, synchronously
        DeviceGuard guard(a->device);
        const size_t sb = static_cast<size_t>(a->total) * sizeof(float);
        cudaError_t ez = cudaMemset(a->exp_avg, 0, sb);
        if (ez == cudaSuccess) ez = cudaMemset(a->exp_avg_sq, 0, sb);
        if (ez == cudaSuccess) ez = cudaDeviceSynchronize();
        if (ez != cudaSuccess) return tfail("hd_adam_state: %s", cudaGetErrorString(ez));
        a->zeroed = true;
    }
    int64_t off = 0;
    for (int i = 0; i < index; ++i) off += ((a->numels[i] + 3) / 4) * 4;
    if (exp_avg) *exp_avg = a->exp_avg + off;
    if (exp_avg_sq) *exp_avg_sq = a->exp_avg_sq + off;
    if (step) *step = a->step;
    return 0;
}

int hd_adam_set_step(hd_adam* a, int64_t step) {
    if (!a || step < 0) return tfail("hd_adam_set_step: bad argument");
    a->step = step;
    return 0;
}

void hd_adam_destroy(hd_adam* a) {
    if (!a) return;
    DeviceGuard guard(a->device);
    if (a->staging) cudaFreeHost(a->staging);
    if (a->table) cudaFree(a->table);
    if (a->exp_avg) cudaFree(a->exp_avg);
    if (a->exp_avg_sq) cudaFree(a->exp_avg_sq);
    delete a;
}

}  // extern "C"
