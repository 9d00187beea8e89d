// Host-side check (g++, no CUDA): the product's Additive/Saw.k / Square.k voice (klang_b200/csrc/kb_graphs.cuh).  For both graphs a
// note is started, rendered over ragged blocks, re-started at another pitch (the partials keep their phase) and rendered again —
// once with the per-tick form (kb_add_tick, what kb_voice_kernel runs) and once with the time-parallel form (kb_add_at per sample in
// any order + kb_add_block_end, what kb_additive_kernel / kb_additive_advance_kernel run).  The two must agree bit for bit, samples
// and state (exit code 1 otherwise); the samples go to stdout as raw float32 and tests/test_host_logic.py compares them with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main() {
	int bad = 0;
	const int blocks[6] = { 700, 1, 129, 1000, 512, 658 };                        // 3000 samples; the second note starts after block 2
	for (int graph : { KB_SY_ADDITIVE_SAW, KB_SY_ADDITIVE_SQUARE, KB_SY_ADDITIVE_NYQUIST }) {
		const KbFs fs = kb_make_fs(graph == KB_SY_ADDITIVE_SAW ? 48000.f : 44100.f);
		KbAddVoice a, b;
		memset(&a, 0, sizeof(a)); memset(&b, 0, sizeof(b));
		kb_add_construct(graph, a); kb_add_construct(graph, b);
		kb_add_on(fs, a, 57.f); kb_add_on(fs, b, 57.f);
		for (int k = 0; k < 6; k++) {
			if (k == 3) { kb_add_on(fs, a, 88.f); kb_add_on(fs, b, 88.f); }        // high pitch: Square.k drops the partials above Nyquist
			const int n = blocks[k];
			std::vector<float> ya(n), yb(n);
			for (int t = 0; t < n; t++) ya[t] = kb_add_tick(fs, a);
			for (int t = n - 1; t >= 0; t--) yb[t] = kb_add_at(fs, b, (uint32_t)t);
			kb_add_block_end(fs, b, (uint32_t)n);
			if (memcmp(ya.data(), yb.data(), sizeof(float) * n) != 0 || memcmp(&a, &b, sizeof(a)) != 0) bad++;
			fwrite(yb.data(), sizeof(float), n, stdout);
		}
	}
	return bad ? 1 : 0;
}
