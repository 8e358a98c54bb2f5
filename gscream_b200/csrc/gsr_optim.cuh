// gsr_optim.cuh — declarations of the multi-tensor Adam step (gsr_optim.cu), shared with gsr_api.cu.
#pragma once
#include "gsr_common.cuh"

namespace gsr {

constexpr int kAdamMaxTensors = 24;   // tensors per launch (the table travels in kernel-parameter space)

struct AdamTensor {
	float *param;
	const float *grad;
	float *exp_avg, *exp_avg_sq;
	int64_t n;
	float one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, step_size, bias_correction2_sqrt;
	int pad;
};

// tensors: HOST array; all launches asynchronous on `stream`.
cudaError_t launch_adam_step(int n_tensors, const AdamTensor *tensors, cudaStream_t stream);

} // namespace gsr
