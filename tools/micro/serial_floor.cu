// Micro-benchmark (not part of the product): dependent-issue floor of the serial recurrences on one warp of one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -I../../klang_b200/csrc serial_floor.cu -o serial_floor
#include <cstdio>
#include <cuda_runtime.h>
#include "kb_fx_parallel.cuh"

__global__ void k_fadd_chain(float* out, long long* cyc, float c, int iters) {
	float x = out[threadIdx.x];
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int j = 0; j < 64; j++) x = x + c;
	}
	long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_mixed_chain(float* out, long long* cyc, float c, float m, int iters) {
	float x = out[threadIdx.x];
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int j = 0; j < 32; j++) { x = x + c; x = x * m; }
	}
	long long t1 = clock64();
	out[threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the row filter of the Reverb pipeline, `lanes` lanes active, n samples from shared memory
__global__ void k_biquad_row(float* out, long long* cyc, int n, int lanes, int reps) {
	__shared__ __align__(16) float xr[8][512 + 16], yr[8][512 + 16];
	for (int i = threadIdx.x; i < 8 * (512 + 16); i += blockDim.x) (&xr[0][0])[i] = 0.001f * (i % 97);
	__syncthreads();
	float z0 = 0.f, z1 = 0.f;
	const float b0 = 0.2f, b1 = 0.4f, b2 = 0.2f, a1 = -0.5f, a2 = 0.3f;
	long long t0 = clock64();
	if (threadIdx.x < lanes)
		for (int r = 0; r < reps; r++) kb_rv2_filter_row(xr[threadIdx.x & 7], yr[threadIdx.x & 7], n, b0, b1, b2, a1, a2, z0, z1);
	long long t1 = clock64();
	if (threadIdx.x < lanes) out[threadIdx.x] = z0 + z1 + yr[threadIdx.x & 7][n - 1];
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// the same recurrence from registers only (no shared memory): the pure arithmetic chain
__global__ void k_biquad_regs(float* out, long long* cyc, int n) {
	float z0 = out[threadIdx.x], z1 = 0.f, x = 0.37f, acc = 0.f;
	const float b0 = 0.2f, b1 = 0.4f, b2 = 0.2f, a1 = -0.5f, a2 = 0.3f;
	long long t0 = clock64();
	#pragma unroll 8
	for (int i = 0; i < n; i++) {
		const float y = b0 * x + z0;
		z0 = b1 * x - a1 * y + z1;
		z1 = b2 * x - a2 * y;
		acc += y;
	}
	long long t1 = clock64();
	out[threadIdx.x] = z0 + z1 + acc;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_env_run(float* out, long long* cyc, int n, int reps, KbFs fs) {
	__shared__ float row[32][129];
	__shared__ float px[16], py[16];
	if (threadIdx.x < 16) { px[threadIdx.x] = threadIdx.x * 100.f; py[threadIdx.x] = (threadIdx.x & 1) ? 1000.f : 1.f; }
	__syncthreads();
	KbEnvR e; e.r_out = 1.f + threadIdx.x; e.r_target = (threadIdx.x & 1) ? 1e9f : -1e9f; e.r_rate = 0.001f; e.time = 0.f; e.timeInc = 1.f / 48000.f; e.out = 0.f;
	e.r_active = 1; e.stage = KB_ENV_SUSTAIN; e.point = 0; e.loop_start = -1; e.loop_end = -1; e.npoints = 3;
	long long t0 = clock64();
	for (int r = 0; r < reps; r++) kb_envr_run(fs, e, px, py, row[threadIdx.x], n);
	long long t1 = clock64();
	out[threadIdx.x] = e.r_out + row[threadIdx.x][n - 1];
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// the row filter on warp 0 (lanes 0..7) while the warps of the OTHER sub-partitions run a background load:
// mode 0 exit, 1 wait at a named barrier, 2 dependent FP32 chain, 3 shared-memory traffic, 4 global loads (L2 hits), 5 all of 2-4 mixed
__global__ void k_contend(float* out, long long* cyc, const float* g, int mode, int reps) {
	__shared__ __align__(16) float xr[8][512 + 16], yr[8][512 + 16];
	__shared__ float scratch[2048];
	__shared__ volatile int stop;
	for (int i = threadIdx.x; i < 8 * (512 + 16); i += blockDim.x) (&xr[0][0])[i] = 0.001f * (i % 97);
	for (int i = threadIdx.x; i < 2048; i += blockDim.x) scratch[i] = 1.f;
	if (threadIdx.x == 0) stop = 0;
	__syncthreads();
	const int warp = threadIdx.x >> 5;
	if (warp == 0) {
		float z0 = 0.f, z1 = 0.f;
		const float b0 = 0.2f, b1 = 0.4f, b2 = 0.2f, a1 = -0.5f, a2 = 0.3f;
		long long t0 = clock64();
		if (threadIdx.x < 8)
			for (int r = 0; r < reps; r++) kb_rv2_filter_row(xr[threadIdx.x], yr[threadIdx.x], 512, b0, b1, b2, a1, a2, z0, z1);
		long long t1 = clock64();
		if (threadIdx.x < 8) out[threadIdx.x] = z0 + z1 + yr[threadIdx.x][511];
		if (threadIdx.x == 0) { cyc[0] = t1 - t0; stop = 1; }
		__syncwarp();
		if (mode == 1) asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory");
		return;
	}
	if (mode == 0) return;
	if (mode == 1 || (warp & 3) == 0) { if (mode == 1) asm volatile("bar.sync 1, %0;" :: "r"((int)blockDim.x) : "memory"); return; }
	float a = threadIdx.x, b = 1.0001f, c = 0.5f, d = 2.f;
	int it = 0;
	while (!stop) {
		const int m = mode == 5 ? 2 + (warp % 3) : mode;
		if (m == 2) {
			#pragma unroll
			for (int j = 0; j < 32; j++) { a = a * b + c; c = c * b + d; }
		} else if (m == 3) {
			#pragma unroll
			for (int j = 0; j < 8; j++) { a += scratch[(threadIdx.x * 5 + j * 97 + it) & 2047]; scratch[(threadIdx.x + j * 131 + it) & 2047] = a; }
		} else {
			#pragma unroll
			for (int j = 0; j < 8; j++) a += __ldcg(g + ((threadIdx.x * 33 + j * 4099 + it * 7) & 0xfffff));
		}
		it++;
	}
	out[threadIdx.x] = a + c;
}

int main() {
	float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64); cudaMemset(d, 0, 4096);
	long long h;
	for (int rep = 0; rep < 2; rep++) {
		k_fadd_chain<<<1, 32>>>(d, c, 1e-9f, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		if (rep) printf("fadd chain: %.2f cycles/op\n", h / 64000.0);
		k_mixed_chain<<<1, 32>>>(d, c, 1e-9f, 0.999f, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		if (rep) printf("fadd+fmul chain: %.2f cycles/op\n", h / 64000.0);
		k_biquad_regs<<<1, 32>>>(d, c, 65536); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		if (rep) printf("biquad TDF-II from registers: %.2f cycles/sample\n", h / 65536.0);
		for (int lanes : {1, 8, 32}) {
			k_biquad_row<<<1, 32>>>(d, c, 512, lanes, 64); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
			if (rep) printf("biquad row filter (smem, groups of 8), %2d lanes: %.2f cycles/sample\n", lanes, h / (512.0 * 64));
		}
		k_biquad_row<<<1, 256>>>(d, c, 512, 8, 64); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		if (rep) printf("biquad row filter, 8 lanes, 7 idle warps resident: %.2f cycles/sample\n", h / (512.0 * 64));
		if (rep) {
			float* g; cudaMalloc(&g, 4 << 20); cudaMemset(g, 0, 4 << 20);
			const char* names[6] = { "others exit", "others wait at a named barrier", "others: dependent FP32 chains", "others: shared-memory traffic", "others: global loads (L2)", "others: mixed" };
			for (int mode = 0; mode < 6; mode++) {
				k_contend<<<1, 608>>>(d, c, g, mode, 64); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
				printf("row filter on warp 0, 608-thread CTA, %-34s: %.2f cycles/sample\n", names[mode], h / (512.0 * 64));
			}
			cudaFree(g);
		}
		k_env_run<<<1, 32>>>(d, c, 128, 256, kb_make_fs(48000.f)); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		if (rep) printf("envelope run (ramp, 32 lanes up/down mixed): %.2f cycles/tick\n", h / (128.0 * 256));
	}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
