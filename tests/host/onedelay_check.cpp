// Host-side check (g++, no CUDA): the product's Delay/Echo.k and Delay/Feedback.k frame functions (kb_echo_frame / kb_feedback_frame of
// klang_b200/csrc/kb_graphs.cuh, what kb_fx_seq_kernel runs per lane) over a host ring.
// Usage: onedelay_check <graph> <fs> <total> <block> <input.f32> [<block> <ctl> <value>]...   Output to stdout; tests/test_host_logic.py
// compares it with the golden vectors of the compiled reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main(int argc, char** argv) {
	if (argc < 6) return 2;
	const int graph = atoi(argv[1]);
	const KbFs fs = kb_make_fs((float)atof(argv[2]));
	const int total = atoi(argv[3]), block = atoi(argv[4]);
	std::vector<float> io(total), ring(KB_ONEDELAY_RING_FLOATS, 0.f);
	FILE* f = fopen(argv[5], "rb");
	if (!f || fread(io.data(), 4, io.size(), f) != io.size()) return 2;
	fclose(f);
	KbFxHdr h;
	memset(&h, 0, sizeof(h));
	h.controls[0] = kb_dial(0.f, 1.f, 0.5f); h.controls[1] = kb_dial(0.f, 1.f, 0.5f);
	KbOneDelayFx s;
	kb_delay_construct(s.delay, 192000, 0);
	KbIirFx iir = { 0.f };                                                    // Filtering/IIR.k and WahWah.k ride along: same block / event driver
	KbModDelayFx md;
	kb_delay_construct(md.delay, 192000, 0);
	for (int k = 0; k < 3; k++) kb_fsine_init(md.lfo[k]);
	kb_osm_construct(md.tri, 0, 1.0f);
	if (graph == KB_FX_FLANGER) { h.controls[0] = kb_dial(0.1f, 1.0f, 0.75f); h.controls[1] = kb_dial(0.1f, 5.0f, 1.5f); }
	if (graph == KB_FX_MODDELAY) { h.controls[0] = kb_dial(1.f, 10.f, 6.f); h.controls[1] = kb_dial(0.f, 1.f, 0.2f); }
	if (graph == KB_FX_MOD_CHORUS) { h.controls[0] = kb_dial(1.f, 10.f, 6.f); h.controls[1] = kb_dial(0.f, 1.f, 0.1f); }
	KbWahWahFx wah;
	kb_biquad_construct(wah.lpf, KB_BQ_LPF); kb_fsine_init(wah.lfo);
	if (graph == KB_FX_WAHWAH) { h.controls[0] = kb_dial(10.f, 10000.f, 1000.f); h.controls[1] = kb_dial(0.1f, 10.f, 1.f); h.controls[2] = kb_dial(4.f, 10.f, 6.f); }
	for (int b = 0; b * block < total; b++) {
		for (int a = 6; a + 2 < argc; a += 3) if (atoi(argv[a]) == b) kb_control_set(h.controls[atoi(argv[a + 1])], (float)atof(argv[a + 2]));
		const int n = total - b * block < block ? total - b * block : block;
		for (int t = 0; t < n; t++) {
			float& x = io[(size_t)b * block + t];
			x = (graph >= KB_FX_FLANGER && graph <= KB_FX_MOD_CHORUS) ? kb_moddelay_frame(graph, fs, h, md, ring.data(), x) : graph == KB_FX_WAHWAH ? kb_wahwah_frame(fs, h, wah, x) : graph == KB_FX_IIR ? kb_iir_frame(h, iir, x) : graph == KB_FX_ECHO ? kb_echo_frame(fs, h, s, ring.data(), x) : kb_feedback_frame(fs, h, s, ring.data(), x);
		}
	}
	fwrite(io.data(), 4, io.size(), stdout);
	return 0;
}
