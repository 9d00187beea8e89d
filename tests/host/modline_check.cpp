// Host-side check (g++, no CUDA): Flanger.k / Modulation/Chorus.k in their time-parallel form — kb_modline_begin, a write sweep that stashes
// what it overwrites (kb_modline_write_at), a read sweep in any order (kb_modline_read_at: closed-form LFOs, stash-aware taps),
// kb_modline_end — against the frame-sequential kb_moddelay_frame: samples, ring contents and the whole state, bit for bit, over
// ragged blocks with the ring wrapping, control changes, and inputs that contain negative zeros.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main() {
	unsigned seed = 31337u;
	auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
	long long frames = 0, bad = 0, state_bad = 0, zero_delay_frames = 0;
	for (int trial = 0; trial < 12; trial++) {
		const int graph = trial % 3 == 0 ? KB_FX_FLANGER : trial % 3 == 1 ? KB_FX_MOD_CHORUS : KB_FX_MODDELAY;
		const KbFs fs = kb_make_fs(trial % 4 < 2 ? 48000.f : 44100.f);
		KbFxHdr h; memset(&h, 0, sizeof(h));
		if (graph == KB_FX_FLANGER) { h.controls[0] = kb_dial(0.1f, 1.0f, 0.75f); h.controls[1] = kb_dial(0.1f, 5.0f, 1.5f); }
		else if (graph == KB_FX_MODDELAY) { h.controls[0] = kb_dial(1.f, 10.f, 6.f); h.controls[1] = kb_dial(0.f, 1.f, 0.2f); }
		else { h.controls[0] = kb_dial(1.f, 10.f, 6.f); h.controls[1] = kb_dial(0.f, 1.f, 0.1f); }
		KbModDelayFx a;
		memset(&a, 0, sizeof(a));
		kb_delay_construct(a.delay, 192000, 0);
		for (int k = 0; k < 3; k++) kb_fsine_init(a.lfo[k]);
		kb_osm_construct(a.tri, 0, 1.0f);
		if (trial >= 4) a.delay.position = 192000 - 2500;                       // the ring wraps inside the run
		KbModDelayFx b = a;
		std::vector<float> ra(KB_ONEDELAY_RING_FLOATS, 0.f), rb(KB_ONEDELAY_RING_FLOATS, 0.f), old(16384), depth(16384);
		KbFxHdr ha = h, hb = h;                                                  // ModDelay.k's smoother lives in the control block: carried per path
		for (int k = 0; k < 14; k++) {
			const int sizes[7] = { 1, 7, 1000, 1024, 4096, 16384, 9999 };
			const int n = sizes[rnd() % 7];
			if (k % 3 == 1 && graph == KB_FX_FLANGER) {
				const float r = 0.1f + (rnd() % 90) * 0.01f, dp = 0.1f + (rnd() % 490) * 0.01f;
				kb_control_set(ha.controls[0], r); kb_control_set(hb.controls[0], r);
				kb_control_set(ha.controls[1], dp); kb_control_set(hb.controls[1], dp);
			}
			if (k % 3 == 1 && graph == KB_FX_MODDELAY) {
				const float r = 1.f + (rnd() % 900) * 0.01f, dp = (rnd() % 100) * 0.01f;
				kb_control_set(ha.controls[0], r); kb_control_set(hb.controls[0], r);
				kb_control_set(ha.controls[1], dp); kb_control_set(hb.controls[1], dp);
			}
			std::vector<float> x(n), ya(n), yb(n);
			for (int t = 0; t < n; t++) { x[t] = ((int)(rnd() % 20001) - 10000) * 1e-4f; if (rnd() % 50 == 0) x[t] = -0.0f; }
			for (int t = 0; t < n; t++) ya[t] = kb_moddelay_frame(graph, fs, ha, a, ra.data(), x[t]);
			kb_modline_begin(graph, fs, hb, b, depth.data(), n);
			for (int t = n - 1; t >= 0; t--) kb_modline_write_at(b.delay, rb.data(), old.data(), t, x[t]);
			for (int t = n - 1; t >= 0; t--) yb[t] = kb_modline_read_at(graph, fs, hb, b, rb.data(), old.data(), depth.data(), n, t, x[t]);
			if (graph == KB_FX_FLANGER) {
				const float dp = hb.controls[1].value / 1000.f;
				for (int t = 0; t < n; t++) if (kb_osm_at(b.tri, (uint32_t)t) * dp + dp == 0.f) zero_delay_frames++;
			}
			if (memcmp(&ha, &hb, sizeof(ha)) != 0) state_bad++;
			kb_modline_end(graph, b, n);
			frames += n;
			if (memcmp(ya.data(), yb.data(), sizeof(float) * n) != 0) bad++;
			if (memcmp(&a, &b, sizeof(a)) != 0) state_bad++;
		}
		if (memcmp(ra.data(), rb.data(), sizeof(float) * ra.size()) != 0) state_bad++;
	}
	printf("modulated-line time-parallel form: %lld frames (%lld with a zero delay), %lld block mismatches, %lld state / ring mismatches\n",
	       frames, zero_delay_frames, bad, state_bad);
	return (bad || state_bad) ? 1 : 0;
}
