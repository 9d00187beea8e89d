// Host-side check (g++, no CUDA): Echo.k's time-parallel form (kb_echo_write_at for every frame of the block, then kb_echo_read_at for every
// frame in any order, then the position advance — what kb_echo_write_kernel / kb_echo_read_kernel / kb_onedelay_advance_kernel run)
// against the frame-sequential kb_echo_frame: samples, ring contents and position, bit for bit, over ragged blocks with the ring
// wrapping around, moving and fractional delay times.  Blocks for which kb_echo_parallel_ok says no run sequentially, as in the library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

int main() {
	unsigned seed = 4242u;
	auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
	long long frames = 0, bad = 0, ring_bad = 0, par_blocks = 0, seq_blocks = 0;
	for (int trial = 0; trial < 6; trial++) {
		const KbFs fs = kb_make_fs(trial % 2 ? 44100.f : 48000.f);
		KbFxHdr h; memset(&h, 0, sizeof(h));
		h.controls[0] = kb_dial(0.f, 1.f, 0.5f); h.controls[1] = kb_dial(0.f, 1.f, 0.5f);
		KbOneDelayFx a, b;
		kb_delay_construct(a.delay, 192000, 0); b = a;
		std::vector<float> ra(KB_ONEDELAY_RING_FLOATS, 0.f), rb(KB_ONEDELAY_RING_FLOATS, 0.f);
		if (trial >= 3) { a.delay.position = b.delay.position = 192000 - 5000; }          // the ring wraps inside the run
		for (int k = 0; k < 14; k++) {
			const int sizes[7] = { 1, 7, 1000, 1024, 4096, 16384, 9999 };
			const int n = sizes[rnd() % 7];
			if (k % 3 == 0) {
				const float choices[6] = { 0.01f, 0.0123f, 0.5f, 1.0f, 0.0f, 0.00001f };      // 0 and 1e-5: delay below one frame -> sequential
				kb_control_set(h.controls[0], choices[rnd() % 6]);
				kb_control_set(h.controls[1], (rnd() % 100) * 0.01f);
			}
			std::vector<float> x(n), ya(n), yb(n);
			for (int t = 0; t < n; t++) x[t] = ((int)(rnd() % 20001) - 10000) * 1e-4f;
			for (int t = 0; t < n; t++) ya[t] = kb_echo_frame(fs, h, a, ra.data(), x[t]);
			if (kb_echo_parallel_ok(fs, n, h.controls[0].value)) {
				par_blocks++;
				for (int t = n - 1; t >= 0; t--) kb_echo_write_at(b, rb.data(), t, x[t]);
				for (int t = n - 1; t >= 0; t--) yb[t] = kb_echo_read_at(fs, h, b, rb.data(), t, x[t]);
				b.delay.position = (b.delay.position + n) % b.delay.SIZE;
			} else {
				seq_blocks++;
				for (int t = 0; t < n; t++) yb[t] = kb_echo_frame(fs, h, b, rb.data(), x[t]);
			}
			frames += n;
			if (memcmp(ya.data(), yb.data(), sizeof(float) * n) != 0) bad++;
			if (a.delay.position != b.delay.position) ring_bad++;
		}
		if (memcmp(ra.data(), rb.data(), sizeof(float) * ra.size()) != 0) ring_bad++;
	}
	printf("echo time-parallel form: %lld frames, %lld blocks parallel, %lld sequential, %lld block mismatches, %lld ring / position mismatches\n",
	       frames, par_blocks, seq_blocks, bad, ring_bad);
	return (bad || ring_bad || par_blocks < 20 || seq_blocks < 3) ? 1 : 0;
}
