// gsr_optim.cu — multi-tensor Adam step, SURVEY.md section 8f rank 4 (second half).
//
// Replaces `gaussians.optimizer.step()` (train.py:611) for the optimizer GaussianModel.training_setup builds with
// `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over 7 per-anchor tensors + 16 MLP weight / bias tensors
// (scene/gaussian_model.py:374-407).  torch's foreach implementation issues ~10 multi-tensor launches that each re-read or re-write
// the parameter / gradient / moment arrays (7 reads + 5 writes of every element instead of 4 + 3); here ONE launch covers every
// tensor of every parameter group: a block finds its tensor in a table passed by value in kernel-parameter space, then streams
// param / grad / exp_avg / exp_avg_sq once with 128-bit accesses.  HBM-bound: 28 B per element.
//
// Arithmetic follows torch.optim.Adam (amsgrad = False, maximize = False), in its order of operations:
//     grad    += weight_decay * param                       (only when weight_decay != 0)
//     exp_avg  = exp_avg + (grad - exp_avg) * (1 - beta1)                      (lerp_)
//     exp_avg_sq = exp_avg_sq * beta2 + (1 - beta2) * grad * grad              (mul_, addcmul_)
//     denom    = sqrt(exp_avg_sq) / sqrt(1 - beta2^step) + eps
//     param    = param - (lr / (1 - beta1^step)) * (exp_avg / denom)           (addcdiv_)
// The two bias corrections are computed on the host in double precision from each tensor's own step count, as torch does.
#include "gsr_common.cuh"
#include "gsr_optim.cuh"

namespace gsr {

namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamVec = 4;                                   // elements per thread per iteration (one float4)
constexpr int kAdamIters = 4;                                 // iterations per block
constexpr int kAdamChunk = kAdamThreads * kAdamVec * kAdamIters; // 4096 elements per block

struct AdamTable {
	AdamTensor t[kAdamMaxTensors];
	int block_end[kAdamMaxTensors];                           // exclusive prefix of blocks per tensor
	int n;
};

__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, const AdamTensor &t)
{
	if (t.weight_decay != 0.f) g = fmaf(t.weight_decay, p, g);
	m = m + (g - m) * t.one_minus_beta1;
	v = v * t.beta2 + t.one_minus_beta2 * g * g;
	const float denom = sqrtf(v) / t.bias_correction2_sqrt + t.eps;
	p = p - t.step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const __grid_constant__ AdamTable tab)
{
	int k = 0;
	while (k < tab.n - 1 && (int)blockIdx.x >= tab.block_end[k]) k++;
	const AdamTensor &t = tab.t[k];
	const int first_block = k == 0 ? 0 : tab.block_end[k - 1];
	const int64_t base = (int64_t)((int)blockIdx.x - first_block) * kAdamChunk;
	const bool vec_ok = ((((uintptr_t)t.param | (uintptr_t)t.grad | (uintptr_t)t.exp_avg | (uintptr_t)t.exp_avg_sq) & 15) == 0);
#pragma unroll
	for (int it = 0; it < kAdamIters; it++) {
		const int64_t i = base + ((int64_t)it * kAdamThreads + threadIdx.x) * kAdamVec;
		if (i >= t.n) break;
		if (vec_ok && i + kAdamVec <= t.n) {
			float4 p = *reinterpret_cast<float4 *>(t.param + i);
			const float4 g = *reinterpret_cast<const float4 *>(t.grad + i);
			float4 m = *reinterpret_cast<float4 *>(t.exp_avg + i);
			float4 v = *reinterpret_cast<float4 *>(t.exp_avg_sq + i);
			adam_update(p.x, g.x, m.x, v.x, t);
			adam_update(p.y, g.y, m.y, v.y, t);
			adam_update(p.z, g.z, m.z, v.z, t);
			adam_update(p.w, g.w, m.w, v.w, t);
			*reinterpret_cast<float4 *>(t.param + i) = p;
			*reinterpret_cast<float4 *>(t.exp_avg + i) = m;
			*reinterpret_cast<float4 *>(t.exp_avg_sq + i) = v;
		} else {
			for (int64_t j = i; j < i + kAdamVec && j < t.n; j++) {
				float p = t.param[j], m = t.exp_avg[j], v = t.exp_avg_sq[j];
				adam_update(p, t.grad[j], m, v, t);
				t.param[j] = p; t.exp_avg[j] = m; t.exp_avg_sq[j] = v;
			}
		}
	}
}

} // namespace

cudaError_t launch_adam_step(int n_tensors, const AdamTensor *tensors, cudaStream_t stream)
{
	for (int first = 0; first < n_tensors;) {
		AdamTable tab;
		tab.n = 0;
		int blocks = 0;
		int k = first;
		for (; k < n_tensors && tab.n < kAdamMaxTensors; k++) {
			if (tensors[k].n <= 0) continue;
			const int64_t nb = (tensors[k].n + kAdamChunk - 1) / kAdamChunk;
			if (nb > (int64_t)0x7fffffff - blocks) return cudaErrorInvalidValue;
			tab.t[tab.n] = tensors[k];
			blocks += (int)nb;
			tab.block_end[tab.n] = blocks;
			tab.n++;
		}
		first = k; // (not first + kAdamMaxTensors: skipped empty descriptors do not count against the table)
		if (tab.n == 0) continue;
		adam_step_kernel<<<blocks, kAdamThreads, 0, stream>>>(tab);
		count_launch();
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) return e;
	}
	return cudaSuccess;
}

} // namespace gsr
