// Micro-probe (not product code): issue rate of legacy mma.sync m16n8k8 TF32 on sm_100a, per SM, as a function of
// resident warps.  Informs whether the blend kernels' per-warp 32x32 products could go to the tensor pipe (3xTF32).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(float *out, int iters)
{
	float c[4][4] = {};
	unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f900000u, 0x3fa00000u, 0x3fb00000u}, b[2] = {0x3f800000u, 0x3f880000u + threadIdx.x};
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 4; k++)
			asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
			             : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
			             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
	}
	float s = 0;
	for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) s += c[k][j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void probe_ffma(float *out, int iters)
{
	float c[8] = {1, 2, 3, 4, 5, 6, 7, 8};
	float a = 1.0001f + threadIdx.x * 1e-7f, b = 0.5f;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 8; k++) c[k] = fmaf(c[k], a, b);
	}
	float s = 0;
	for (int k = 0; k < 8; k++) s += c[k];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int sms = p.multiProcessorCount;
	float *out;
	cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	printf("device %s SMs %d clock attr %d kHz\n", p.name, sms, clk_khz);
	const int iters = 20000;
	for (int warps = 4; warps <= 32; warps *= 2) {
		const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
		const int blocks = sms * ((warps * 32 + threads - 1) / threads);
		for (int which = 0; which < 2; which++) {
			for (int rep = 0; rep < 2; rep++) {
				cudaEventRecord(e0);
				if (which == 0) probe<<<blocks, threads>>>(out, iters);
				else probe_ffma<<<blocks, threads>>>(out, iters);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
			}
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			const double n = (double)iters * (which == 0 ? 4 : 8) * warps; // warp-instructions per SM
			printf("%s warps/SM %2d: %.3f ms, %.2f warp-instr/us/SM (%.3f per clk at 1.965 GHz)%s\n", which == 0 ? "mma.m16n8k8.tf32" : "ffma            ",
			       warps, ms, n / (ms * 1e3), n / (ms * 1e3) / 1965.0, which == 0 ? "" : "");
		}
	}
	printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
