// Micro-benchmark (not part of the product): what the envelope warp of the C2 kernel pays per tick — the two dependent FADD chains, the
// exit test, the stores — one warp alone on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -ftz=false -O3 -I../../klang_b200/csrc env_floor.cu -o env_floor
#include <cstdio>
#include <cuda_runtime.h>
#include "kb_prims.cuh"

struct Rows { alignas(16) float r[32][132]; };

template <int MODE>   // 0 kb_envr_run, 1 kb_envr_run16<true>, 2 kb_envr_run16<false>
__global__ void k_env(float* out, long long* cyc, int n, int reps, int lanes, KbFs fs) {
	__shared__ Rows rows;
	__shared__ float px[16], py[16];
	if (threadIdx.x < 16) { px[threadIdx.x] = threadIdx.x * 100.f; py[threadIdx.x] = (threadIdx.x & 1) ? 1000.f : 1.f; }
	__syncthreads();
	KbEnvR e; e.r_out = 1.f + threadIdx.x; e.r_target = (threadIdx.x & 1) ? 1e9f : -1e9f; e.r_rate = 0.001f; e.time = 0.f; e.timeInc = 1.f / 48000.f; e.out = 0.f;
	e.r_active = 1; e.stage = KB_ENV_SUSTAIN; e.point = 0; e.loop_start = -1; e.loop_end = -1; e.npoints = 3;
	long long t0 = clock64();
	if (threadIdx.x < lanes)
		for (int r = 0; r < reps; r++) {
			if (MODE == 0) kb_envr_run(fs, e, px, py, rows.r[threadIdx.x], n);
			else if (MODE == 1) kb_envr_run16<true>(fs, e, px, py, rows.r[threadIdx.x], n);
			else if (MODE == 2) kb_envr_run16<false>(fs, e, px, py, rows.r[threadIdx.x], n);
			else kb_envr_run_tile<true>(fs, e, px, py, rows.r[threadIdx.x], n);
		}
	long long t1 = clock64();
	out[threadIdx.x] = e.r_out + rows.r[threadIdx.x][n - 1];
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// hand-written group loops: FLAGS bit 0 = time chain, 1 = exit test, 2 = 128-bit stores, 4 = scalar stores
template <int FLAGS, int GROUP>
__global__ void k_loop(float* out, long long* cyc, int n, int reps, int lanes, float srate, float tinc, float hi) {
	__shared__ Rows rows;
	float r = 1.f + threadIdx.x, time = 0.f;
	float* row = rows.r[threadIdx.x];
	int broke = 0;
	long long t0 = clock64();
	if (threadIdx.x < lanes)
		for (int rep = 0; rep < reps; rep++) {
			for (int t = 0; t + GROUP <= n; t += GROUP) {
				float rr[GROUP + 1], tt = time;
				rr[0] = r;
				#pragma unroll
				for (int j = 0; j < GROUP; j++) { rr[j + 1] = rr[j] + srate; if (FLAGS & 1) tt = tt + tinc; }
				if (FLAGS & 2) { if (!((rr[GROUP] < hi) & (tt < hi))) { broke++; break; } }
				if (FLAGS & 4) {
					#pragma unroll
					for (int q = 0; q < GROUP / 4; q++) *reinterpret_cast<float4*>(row + t + 4 * q) = make_float4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
				}
				if (FLAGS & 8) {
					#pragma unroll
					for (int j = 0; j < GROUP; j++) row[t + j] = rr[j];
				}
				r = rr[GROUP]; time = tt;
			}
		}
	long long t1 = clock64();
	out[threadIdx.x] = r + time + broke + row[n - 1];
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// issue rate of independent FADD / FMUL / mixed streams from one warp (which pipes take FADD?)
template <int KIND>
__global__ void k_pipe(float* out, long long* cyc, float c, int iters) {
	float x[8];
	for (int j = 0; j < 8; j++) x[j] = out[threadIdx.x] + j;
	long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
		#pragma unroll
		for (int u = 0; u < 8; u++) {
			#pragma unroll
			for (int j = 0; j < 8; j++) {
				if (KIND == 0) x[j] = x[j] + c;
				else if (KIND == 1) x[j] = x[j] * c;
				else x[j] = (j & 1) ? x[j] + c : x[j] * c;
			}
		}
	}
	long long t1 = clock64();
	float a = 0; for (int j = 0; j < 8; j++) a += x[j];
	out[threadIdx.x] = a;
	if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
	float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 64); cudaMemset(d, 0, 4096);
	long long h;
	const int n = 128, reps = 256;
	const KbFs fs = kb_make_fs(48000.f);
#define RUN(NAME, ...) for (int rep = 0; rep < 2; rep++) { __VA_ARGS__; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("%-64s %.2f cycles/tick\n", NAME, h / double(n * reps)); }
	for (int lanes : {32, 14}) {
		printf("-- %d lanes\n", lanes);
		RUN("kb_envr_run (groups of 32, scalar stores)", k_env<0><<<1, 32>>>(d, c, n, reps, lanes, fs));
		RUN("kb_envr_run16<true> (uniform groups of 16, 128-bit stores)", k_env<1><<<1, 32>>>(d, c, n, reps, lanes, fs));
		RUN("kb_envr_run16<false> (uniform groups of 16, scalar stores)", k_env<2><<<1, 32>>>(d, c, n, reps, lanes, fs));
		RUN("kb_envr_run_tile<true> (one test per tile, 128-bit stores)", k_env<3><<<1, 32>>>(d, c, n, reps, lanes, fs));
		RUN("loop16: r chain only", (k_loop<0, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop16: r + time chains", (k_loop<1, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop16: r + time + exit test", (k_loop<3, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop16: r + time + exit test + 128-bit stores", (k_loop<7, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop16: r + time + exit test + scalar stores", (k_loop<11, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop16: r + time + 128-bit stores (no exit test)", (k_loop<5, 16><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop32: r + time + exit test + 128-bit stores", (k_loop<7, 32><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop64: r + time + exit test + 128-bit stores", (k_loop<7, 64><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
		RUN("loop32: r chain + 128-bit stores", (k_loop<4, 32><<<1, 32>>>(d, c, n, reps, lanes, 0.001f, 2e-5f, 1e30f)));
	}
	for (int rep = 0; rep < 2; rep++) {
		k_pipe<0><<<1, 32>>>(d, c, 1.000001f, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("8 independent FADD chains: %.2f cycles/instruction\n", h / 64000.0);
		k_pipe<1><<<1, 32>>>(d, c, 1.000001f, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("8 independent FMUL chains: %.2f cycles/instruction\n", h / 64000.0);
		k_pipe<2><<<1, 32>>>(d, c, 1.000001f, 1000); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("4 FADD + 4 FMUL chains: %.2f cycles/instruction\n", h / 64000.0);
	}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
