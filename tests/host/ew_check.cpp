// Host-side check (g++, no CUDA): the product's elementwise effects (Pan.k, RM.k, Tremolo.k, Clipping.k) as the library runs a block —
// Sine::set(rate) once per block on the mirror (fx_prepare), every sample from kb_ew_sample(t) (kb_elementwise_kernel), then the LFO
// advanced by the block's ticks (kb_lfo_advance_kernel).  Usage: ew_check <graph> <fs> <channels> <total> <block> <input.f32> [<block> <ctl> <value>]...
// The planar input [channels][total] is read from the file, the output written to stdout; tests/test_host_logic.py compares it with
// the golden vectors of the compiled reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main(int argc, char** argv) {
	if (argc < 7) return 2;
	const int graph = atoi(argv[1]);
	const KbFs fs = kb_make_fs((float)atof(argv[2]));
	const int channels = atoi(argv[3]), total = atoi(argv[4]), block = atoi(argv[5]);
	std::vector<float> io((size_t)channels * total);
	FILE* f = fopen(argv[6], "rb");
	if (!f || fread(io.data(), 4, io.size(), f) != io.size()) return 2;
	fclose(f);
	KbControl c[2] = { kb_dial(0.f, 1.f, 0.5f), kb_dial(0.f, 0.5f, 0.5f) };                  // Pan.k:10
	if (graph == KB_FX_RM || graph == KB_FX_TREMOLO) c[0] = kb_dial(1.f, graph == KB_FX_RM ? 1000.f : 10.f, 6.f);
	if (graph == KB_FX_CLIPPING) c[0] = kb_dial(1.f, 11.f, 1.f);
	if (graph == KB_FX_FUNCTIONS) c[0] = kb_dial(1.f, 25.f, 1.f);
	if (graph == KB_FX_MUTE) c[0] = kb_dial(0.f, 1.f, 0.f);
	KbFastSine lfo; kb_fsine_init(lfo);
	for (int b = 0; b * block < total; b++) {
		for (int a = 7; a + 2 < argc; a += 3) if (atoi(argv[a]) == b) kb_control_set(c[atoi(argv[a + 1])], (float)atof(argv[a + 2]));
		if (graph == KB_FX_RM || graph == KB_FX_TREMOLO) kb_fsine_set_f(fs, lfo, c[0].value);
		const int n = total - b * block < block ? total - b * block : block;
		for (int ch = 0; ch < channels; ch++)
			for (int t = n - 1; t >= 0; t--) {                                                // any order: samples are independent
				float& x = io[(size_t)ch * total + (size_t)b * block + t];
				x = kb_ew_sample(graph, c[0].value, c[1].value, lfo, ch, (uint32_t)t, x);
			}
		lfo.position += (uint32_t)n * (uint32_t)lfo.increment;
	}
	fwrite(io.data(), 4, io.size(), stdout);
	return 0;
}
