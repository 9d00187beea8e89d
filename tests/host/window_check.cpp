// Host-side check (g++, no CUDA): the product's Envelope::Follower::Window<64> loop (kb_window_follower_run of
// klang_b200/csrc/kb_prims.cuh — what kb_prim_filter_kernel runs for kinds 15 / 16) on the input found in argv[1]
// (raw: int32 rms, float A, float R, int32 n, in[n]); writes out[n] then coeffs[5] as raw float32 to stdout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_prims.cuh"

int main(int argc, char** argv) {
	if (argc < 2) return 2;
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 2;
	int32_t rms = 0, n = 0;
	float A = 0, R = 0;
	if (fread(&rms, 4, 1, f) != 1 || fread(&A, 4, 1, f) != 1 || fread(&R, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1 || n <= 0) return 2;
	std::vector<float> in(n), out(n);
	if (fread(in.data(), 4, n, f) != (size_t)n) return 2;
	fclose(f);
	float coeffs[5];
	kb_window_follower_run(rms, A, R, n, in.data(), out.data(), coeffs);
	fwrite(out.data(), 4, n, stdout);
	fwrite(coeffs, 4, 5, stdout);
	return 0;
}
