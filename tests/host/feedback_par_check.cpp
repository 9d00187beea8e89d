// Host-side check (g++, no CUDA): Feedback.k's chunk-parallel form (chunks of kb_feedback_chunk frames in order, the frames of a chunk in any
// order through kb_feedback_at, then the position advance — what kb_feedback_par_kernel runs) against the frame-sequential
// kb_feedback_frame: samples, ring contents and position, bit for bit, over ragged blocks with the ring wrapping, integer / fractional
// / moving delays down to the 3-frame limit, and the frame-by-frame path for shorter delays.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main() {
	unsigned seed = 99u;
	auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
	long long frames = 0, bad = 0, ring_bad = 0, chunked = 0, serial = 0;
	for (int trial = 0; trial < 6; trial++) {
		const KbFs fs = kb_make_fs(trial % 2 ? 44100.f : 48000.f);
		KbFxHdr h; memset(&h, 0, sizeof(h));
		h.controls[0] = kb_dial(0.f, 1.f, 0.5f); h.controls[1] = kb_dial(0.f, 1.f, 0.5f);
		KbOneDelayFx a, b;
		kb_delay_construct(a.delay, 192000, 0); b = a;
		std::vector<float> ra(KB_ONEDELAY_RING_FLOATS, 0.f), rb(KB_ONEDELAY_RING_FLOATS, 0.f);
		if (trial >= 3) { a.delay.position = b.delay.position = 192000 - 3000; }
		for (int k = 0; k < 16; k++) {
			const int sizes[7] = { 1, 7, 1000, 1024, 4096, 16384, 9999 };
			const int n = sizes[rnd() % 7];
			if (k % 2 == 0) {
				// 3 / fs and 4 / fs: integer delays at the limit; 0.00008 * fs: 3.5 - 3.8 frames; 0.00001: below the limit
				const float choices[8] = { 0.005f, 0.0123f, 0.5f, 3.0f / fs.f, 4.0f / fs.f, 0.00008f, 0.0f, 0.00001f };
				kb_control_set(h.controls[0], choices[rnd() % 8]);
				kb_control_set(h.controls[1], (rnd() % 95) * 0.01f);
			}
			std::vector<float> x(n), ya(n), yb(n);
			for (int t = 0; t < n; t++) x[t] = ((int)(rnd() % 20001) - 10000) * 1e-4f;
			for (int t = 0; t < n; t++) ya[t] = kb_feedback_frame(fs, h, a, ra.data(), x[t]);
			const int chunk = kb_feedback_chunk(fs, n, h.controls[0].value);
			if (chunk == 0) { serial++; for (int t = 0; t < n; t++) yb[t] = kb_feedback_at(fs, h, b, rb.data(), t, x[t]); }
			else {
				chunked++;
				for (int c0 = 0; c0 < n; c0 += chunk) {
					const int c1 = n < c0 + chunk ? n : c0 + chunk;
					for (int t = c1 - 1; t >= c0; t--) yb[t] = kb_feedback_at(fs, h, b, rb.data(), t, x[t]);
				}
			}
			b.delay.position = (b.delay.position + n) % b.delay.SIZE;
			frames += n;
			if (memcmp(ya.data(), yb.data(), sizeof(float) * n) != 0) bad++;
			if (a.delay.position != b.delay.position) ring_bad++;
		}
		if (memcmp(ra.data(), rb.data(), sizeof(float) * ra.size()) != 0) ring_bad++;
	}
	printf("feedback chunk-parallel form: %lld frames, %lld blocks chunked, %lld frame by frame, %lld block mismatches, %lld ring / position mismatches\n",
	       frames, chunked, serial, bad, ring_bad);
	return (bad || ring_bad || chunked < 30 || serial < 5) ? 1 : 0;
}
