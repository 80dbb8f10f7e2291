// Fit of the transfer head on frozen-trunk embeddings.
//
// Replaces the first phase of `transfer_learning` (SpeakerIdentification/scripts/speaker_identification.py:401-432):
//     sliced_base_model = Model(base_model.input, base_model.layers[-2].output); sliced_base_model.trainable = False
//     outputs = Dense(dim, activation='sigmoid', name='customized_dense')(x)
//     model.compile(loss="categorical_crossentropy", optimizer=RMSprop(lr=0.0001)); model.fit(batch_size=16, epochs=500)
// With the trunk frozen its 512-d outputs (mmla_net_embed) are constants, so the fit is a logistic-type regression on
// [M,512] embeddings.  Keras semantics restated: categorical_crossentropy on non-softmax outputs divides them by their
// sum, clips to [1e-7, 1 - 1e-7] and takes -sum_k y_k log q_k, averaged over the mini-batch; RMSprop keeps
// rms = rho rms + (1 - rho) g^2 and steps by lr g / (sqrt(rms) + eps) (rho 0.9, eps 1e-7, no momentum); the last
// mini-batch of an epoch may be short.  Keras reshuffles the samples every epoch: the visiting order is an input, so a
// run is reproducible and checkable against a CPU fit from the same initial weights.
//
// The steps are strictly sequential (48 000 of them for 1 536 samples x 500 epochs), each a [16,512]x[512,n] product and
// its transpose: one persistent CTA keeps the weights (class-major, padded rows) and the mini-batch in shared memory,
// thread i owns row i of the kernel and its RMS accumulators (registers).
#include "common.cuh"

namespace {

constexpr int kIn = 512, kMaxN = 64, kMaxBatch = 32, kLd = kIn + 1;

struct FitArgs {
    const float* embed;
    const float* y;
    const int32_t* order;
    float* kernel;
    float* bias;
    float* loss_out;
    long long n_samples;
    int n, epochs, batch;
    float lr, rho, eps;
};

__global__ void __launch_bounds__(kIn, 1) head_fit_kernel(const FitArgs a) {
    extern __shared__ float sm[];
    float* wt = sm;                                  // [n][513]   kernel, class-major
    float* xb = wt + kMaxN * kLd;                    // [batch][513]
    float* dz = xb + kMaxBatch * kLd;                // [batch][64]  logits, then dL/dz
    float* bs = dz + kMaxBatch * kMaxN;              // [64] bias
    float* ls = bs + kMaxN;                          // [batch] per-sample loss
    __shared__ int idx[kMaxBatch];
    const int tid = threadIdx.x, n = a.n;
    for (int j = 0; j < n; ++j) wt[j * kLd + tid] = a.kernel[static_cast<long long>(tid) * n + j];
    if (tid < n) bs[tid] = a.bias[tid];
    float v[kMaxN];
#pragma unroll
    for (int j = 0; j < kMaxN; ++j) v[j] = 0.f;
    float vb = 0.f;
    __syncthreads();
    const long long M = a.n_samples;
    for (int ep = 0; ep < a.epochs; ++ep) {
        float ep_loss = 0.f;                         // thread 0 only
        for (long long s0 = 0; s0 < M; s0 += a.batch) {
            const int nb = static_cast<int>(min(static_cast<long long>(a.batch), M - s0));
            if (tid < nb) idx[tid] = a.order[static_cast<long long>(ep) * M + s0 + tid];
            __syncthreads();
            for (int s = 0; s < nb; ++s) xb[s * kLd + tid] = a.embed[static_cast<long long>(idx[s]) * kIn + tid];
            __syncthreads();
            // forward: logits[s][j]
            for (int o = tid; o < nb * n; o += kIn) {
                const int s = o / n, j = o - s * n;
                const float* xr = xb + s * kLd;
                const float* wr = wt + j * kLd;
                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
                for (int i = 0; i < kIn; i += 4) {
                    acc0 = fmaf(xr[i], wr[i], acc0);
                    acc1 = fmaf(xr[i + 1], wr[i + 1], acc1);
                    acc2 = fmaf(xr[i + 2], wr[i + 2], acc2);
                    acc3 = fmaf(xr[i + 3], wr[i + 3], acc3);
                }
                dz[s * kMaxN + j] = (acc0 + acc1) + (acc2 + acc3) + bs[j];
            }
            __syncthreads();
            // per sample: sigmoid, normalise, clip, loss and dL/dz (mean over the mini-batch)
            if (tid < nb) {
                float* z = dz + tid * kMaxN;
                const float* yy = a.y + static_cast<long long>(idx[tid]) * n;
                float S = 0.f;
                for (int j = 0; j < n; ++j) {
                    const float p = 1.f / (1.f + expf(-z[j]));
                    z[j] = p;
                    S += p;
                }
                float loss = 0.f, cross = 0.f;       // cross = sum_k y_k [q_k inside the clip range] p_k / (q_k S^2)
                for (int j = 0; j < n; ++j) {
                    const float q = z[j] / S;
                    const float qc = fminf(fmaxf(q, 1e-7f), 1.f - 1e-7f);
                    loss -= yy[j] * logf(qc);
                    if (q >= 1e-7f && q <= 1.f - 1e-7f) cross += yy[j] * z[j] / (qc * S * S);
                }
                const float inv_nb = 1.f / nb;
                for (int j = 0; j < n; ++j) {
                    const float p = z[j], q = p / S;
                    const float qc = fminf(fmaxf(q, 1e-7f), 1.f - 1e-7f);
                    float dp = cross;
                    if (q >= 1e-7f && q <= 1.f - 1e-7f) dp -= yy[j] / (qc * S);
                    z[j] = dp * p * (1.f - p) * inv_nb;
                }
                ls[tid] = loss;
            }
            __syncthreads();
            // backward + RMSprop: thread i updates row i of the kernel, thread j < n the bias
#pragma unroll
            for (int j = 0; j < kMaxN; ++j) {
                if (j < n) {
                    float g = 0.f;
                    for (int s = 0; s < nb; ++s) g = fmaf(xb[s * kLd + tid], dz[s * kMaxN + j], g);
                    v[j] = a.rho * v[j] + (1.f - a.rho) * g * g;
                    wt[j * kLd + tid] -= a.lr * g / (sqrtf(v[j]) + a.eps);
                }
            }
            if (tid < n) {
                float g = 0.f;
                for (int s = 0; s < nb; ++s) g += dz[s * kMaxN + tid];
                vb = a.rho * vb + (1.f - a.rho) * g * g;
                bs[tid] -= a.lr * g / (sqrtf(vb) + a.eps);
            }
            if (tid == 0)
                for (int s = 0; s < nb; ++s) ep_loss += ls[s];
            __syncthreads();
        }
        if (tid == 0 && a.loss_out) a.loss_out[ep] = ep_loss / static_cast<float>(M);
    }
    for (int j = 0; j < n; ++j) a.kernel[static_cast<long long>(tid) * n + j] = wt[j * kLd + tid];
    if (tid < n) a.bias[tid] = bs[tid];
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int mmla_head_fit(const float* embed, const float* y_onehot, int64_t n_samples,
                                                                    int32_t n_classes, const int32_t* order, int32_t epochs,
                                                                    int32_t batch_size, float lr, float rho, float eps,
                                                                    float* kernel, float* bias, float* loss_out, void* stream) {
    MMLA_REQUIRE(embed && y_onehot && order && kernel && bias, MMLA_EINVAL, "head_fit: null argument");
    MMLA_REQUIRE(n_samples >= 1 && n_classes >= 1 && n_classes <= kMaxN, MMLA_EINVAL, "head_fit: n_classes must be in [1, %d]", kMaxN);
    MMLA_REQUIRE(batch_size >= 1 && batch_size <= kMaxBatch && epochs >= 0, MMLA_EINVAL, "head_fit: batch_size must be in [1, %d]", kMaxBatch);
    MMLA_REQUIRE(mmla_num_sms() > 0, MMLA_ECUDA, "head_fit: no CUDA device");
    if (epochs == 0) return MMLA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (static_cast<size_t>(kMaxN) * kLd + static_cast<size_t>(kMaxBatch) * kLd + kMaxBatch * kMaxN + kMaxN + kMaxBatch) * sizeof(float);
    static MmlaPerDeviceOnce once;
    if (once.first())
        MMLA_CUDA_CHECK(cudaFuncSetAttribute(head_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    FitArgs a;
    a.embed = embed; a.y = y_onehot; a.order = order; a.kernel = kernel; a.bias = bias; a.loss_out = loss_out;
    a.n_samples = n_samples; a.n = n_classes; a.epochs = epochs; a.batch = batch_size; a.lr = lr; a.rho = rho; a.eps = eps;
    head_fit_kernel<<<1, kIn, smem, st>>>(a);
    mmla_count_launch("head_fit_kernel", st);
    MMLA_CUDA_CHECK(cudaGetLastError());
    return MMLA_OK;
}
