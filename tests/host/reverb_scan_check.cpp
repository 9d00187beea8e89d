// Host-side check (g++, no CUDA) of the tolerance-mode filter scan (klang_b200/csrc/kb_scan.cuh, used by kb_reverb3_kernel<1> under
// KB_FX_TOLERANCE): how far does the RE-ASSOCIATED evaluation of Reverb.k's 16 damping low-passes move the output away from the
// reference's sequential fp32 evaluation, in units of the parity bar 1e-5 |r| + 1e-6 peak(r)?
//
//   part 1  one Biquad::LPF on noise: sequential TDF-II ticks (kb_biquad_tick = klang.h:5605-5612) against the scan run lane by lane exactly as
//           the warp runs it (5 ticks per lane, Kogge-Stone over the transition-matrix powers), chunk after chunk, for a ladder of cutoffs.
//   part 2  the whole graph: kb_reverb_frame frame by frame (A, the bit-exact path; the caller may dump it to compare with the compiled
//           reference) against the chunked evaluation the kernel performs (B: per chunk, the ring windows of the 16 lines are read ahead,
//           filtered, and the FDN consumes the filter outputs), on the C4 input — uniform noise — for 64 blocks of 4096 frames at 48 kHz,
//           with every bus audible.  B is run twice: with the sequential filter (must equal A bit for bit: proves the chunked read-ahead is
//           the same computation) and with the scan (the number this file exists for).
// Prints one line per case; exits non-zero if the chunked exact form differs from A or an admitted configuration exceeds the bar.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"
#include "../../klang_b200/csrc/kb_scan.cuh"

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static float noise() {                                    // uniform [-0.5, 0.5)
	g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull;
	return (float)((double)(g_seed >> 40) / (double)(1 << 24) - 0.5);
}

// kb_rv3_scan_chunk, lane by lane
static void scan_chunk_host(const KbRv3ScanCoef& c, const float* w, float frac, int ticks, float* yr, float& z0, float& z1) {
	float x[32][KB_RV3_SCAN_P], e0[32], e1[32];
	int cnt[32];
	for (int lane = 0; lane < 32; lane++) {
		const int t0 = lane * KB_RV3_SCAN_P;
		cnt[lane] = std::max(0, std::min(KB_RV3_SCAN_P, ticks - t0));
		float wa = cnt[lane] > 0 ? w[t0] : 0.f;
		for (int j = 0; j < KB_RV3_SCAN_P; j++) {
			const float wb = j < cnt[lane] ? w[t0 + j + 1] : 0.f;
			x[lane][j] = wa + frac * (wb - wa);
			wa = wb;
		}
		e0[lane] = lane == 0 ? z0 : 0.f; e1[lane] = lane == 0 ? z1 : 0.f;
		kb_rv3_scan_run(c, x[lane], cnt[lane], e0[lane], e1[lane], nullptr);
	}
	for (int j = 0; j < 5; j++) {
		float n0[32], n1[32];
		for (int lane = 0; lane < 32; lane++) {
			n0[lane] = e0[lane]; n1[lane] = e1[lane];
			if (lane >= (1 << j)) {
				const float u0 = e0[lane - (1 << j)], u1 = e1[lane - (1 << j)];
				n0[lane] = kb_fma(c.T[j][0], u0, kb_fma(c.T[j][1], u1, e0[lane]));
				n1[lane] = kb_fma(c.T[j][2], u0, kb_fma(c.T[j][3], u1, e1[lane]));
			}
		}
		memcpy(e0, n0, sizeof(e0)); memcpy(e1, n1, sizeof(e1));
	}
	const int last = (ticks - 1) / KB_RV3_SCAN_P;
	float out0 = z0, out1 = z1;
	for (int lane = 0; lane < 32; lane++) {
		float s0 = lane == 0 ? z0 : e0[lane - 1], s1 = lane == 0 ? z1 : e1[lane - 1];
		float yv[KB_RV3_SCAN_P];
		kb_rv3_scan_run(c, x[lane], cnt[lane], s0, s1, yv);
		for (int j = 0; j < cnt[lane]; j++) yr[lane * KB_RV3_SCAN_P + j] = yv[j];
		if (lane == last) { out0 = s0; out1 = s1; }
	}
	z0 = out0; z1 = out1;
}

static double bar_excess(const std::vector<float>& g, const std::vector<float>& r) {   // max |g - r| / (1e-5 |r| + 1e-6 peak)
	double peak = 0;
	for (float v : r) peak = std::max(peak, (double)fabsf(v));
	double worst = 0;
	for (size_t i = 0; i < r.size(); i++) worst = std::max(worst, fabs((double)g[i] - (double)r[i]) / (1e-5 * fabs((double)r[i]) + 1e-6 * peak));
	return worst;
}

// ---- part 2: Reverb.k, chunked evaluation (what kb_reverb3_kernel does), with the sequential filter or the scan
static int plan_chunk(const KbReverb& rv) {
	int chunk = 80;
	for (int k = 0; k < 2; k++) for (int i = 0; i < 4; i++) for (int grp = 0; grp < 2; grp++) {
		const KbDelay& d = (grp ? rv.late[k] : rv.mid[k]).d[i].delay;
		int lag = d.position - d.last_position; if (lag <= 0) lag += d.SIZE;
		chunk = std::min(chunk, (lag - 2) / 4);
	}
	float tmin = 1e30f;
	for (int r = 0; r < rv.count; r++) tmin = std::min(tmin, rv.times[r]);
	chunk = std::min(chunk, (int)tmin - 3);
	return chunk & ~3;
}
static void reverb_chunked(KbFxHdr& h, KbReverb& rv, float* rings, float* L, float* R, int n, bool scan) {
	const KbControl* c = h.controls;
	const float dry = c[0].value, wet = c[4].value;
	const int Lc = plan_chunk(rv);
	if (Lc < 8) { for (int i = 0; i < n; i++) { float ol, orr; kb_reverb_frame(h, rv, rings, L[i], R[i], ol, orr); L[i] = ol; R[i] = orr; } return; }
	static float yv[2][8][200];                                                // [side][line][tick]
	float* ringl = rings + rv.dl.ring; float* ringr = rings + rv.dr.ring;
	for (int g0 = 0; g0 < n; g0 += Lc) {
		const int len = std::min(Lc, n - g0), ticks = 2 * len;
		for (int side = 0; side < 2; side++) for (int j = 0; j < 8; j++) {
			KbRvFDelay& fd = (j < 4 ? rv.mid[side] : rv.late[side]).d[j & 3];
			const float* ring = rings + fd.delay.ring;
			float w[200];
			for (int t = 0; t <= ticks; t++) w[t] = ring[(fd.delay.last_position + t) % fd.delay.SIZE];     // the window, read AHEAD of the chunk's writes
			if (scan && kb_rv3_scan_admissible(fd.filter)) {
				KbRv3ScanCoef sc; kb_rv3_scan_coef(fd.filter, sc);
				scan_chunk_host(sc, w, fd.delay.last_fraction, ticks, yv[side][j], fd.filter.z0, fd.filter.z1);
			} else {
				for (int t = 0; t < ticks; t++) yv[side][j][t] = kb_biquad_tick(fd.filter, w[t] + fd.delay.last_fraction * (w[t + 1] - w[t]));
			}
		}
		for (int t = 0; t < len; t++) {
			const float inl = L[g0 + t], inr = R[g0 + t];
			const float fl = kb_biquad_tick(rv.hpf[0], kb_biquad_tick(rv.lpf[0], inl));
			const float fr = kb_biquad_tick(rv.hpf[1], kb_biquad_tick(rv.lpf[1], inr));
			kb_delay_write(rv.dl, ringl, fl); kb_delay_write(rv.dr, ringr, fr);
			float r1[2] = { 0, 0 };
			for (int d = 0; d < rv.count; d++) {
				float tl, tr;
				kb_sdelay_tap_f(rv.dl, ringl, ringr, rv.times[d], tl, tr);
				r1[0] += tl * rv.gl[d];
				r1[1] += tr * rv.gr[d];
			}
			float refl[2];
			for (int side = 0; side < 2; side++) {
				float in = r1[side], r2 = 0, r3 = 0;
				for (int stage = 0; stage < 2; stage++) {
					KbRvLate& G = stage ? rv.late[side] : rv.mid[side];
					float dl[4], sum = 0;
					for (int i = 0; i < 4; i++) {                           // first tick: writes the previous frame's feedback, reads, filters
						KbRvFDelay& d = G.d[i];
						kb_delay_write(d.delay, rings + d.delay.ring, d.in);
						d.delay.last_position = (d.delay.last_position + 1) % d.delay.SIZE;
						dl[i] = yv[side][stage * 4 + i][2 * t] * d.gain;
					}
					const float M[4][4] = { { 0, 1, 1,-1 }, {-1, 0,-1, 1 }, {-1, 1, 0,-1 }, { 1,-1, 1, 0 } };
					for (int r = 0; r < 4; r++) G.d[r].in = (M[r][0] * dl[0] + M[r][1] * dl[1] + M[r][2] * dl[2] + M[r][3] * dl[3]) + in;
					for (int i = 0; i < 4; i++) {                           // second tick
						KbRvFDelay& d = G.d[i];
						kb_delay_write(d.delay, rings + d.delay.ring, d.in);
						d.delay.last_position = (d.delay.last_position + 1) % d.delay.SIZE;
						const float o = yv[side][stage * 4 + i][2 * t + 1] * d.gain;
						sum = i == 0 ? o : sum + o;
					}
					if (stage == 0) { r2 = sum; in = sum; } else r3 = sum;
				}
				refl[side] = (r1[side] * c[1].value + r2 * c[2].value) + r3 * c[3].value;
			}
			L[g0 + t] = inl * dry + refl[0] * wet;
			R[g0 + t] = inr * dry + refl[1] * 0.f;
		}
	}
}

int main(int argc, char** argv) {
	const KbFs fs = kb_make_fs(48000.f);
	int fails = 0;
	// ---------------------------------------------------------------------------------------------------- part 1
	printf("part 1: Biquad::LPF (Q = 1/sqrt 2) on uniform noise, 262144 ticks in chunks of 150, fs 48 kHz\n");
	printf("%10s %8s %10s %26s %26s\n", "cutoff Hz", "a2", "admitted", "seq fp32 vs fp64 [bar]", "scan vs seq fp32 [bar]");
	const float cuts[] = { 15000.f, 10000.f, 7000.f, 5000.f, 4000.f, 3000.f, 2000.f, 1000.f, 500.f };
	for (float f : cuts) {
		KbBiquad q; kb_biquad_construct(q, KB_BQ_LPF); kb_biquad_set_f(fs, q, f);
		KbBiquad s = q;
		const int N = 262144;
		std::vector<float> x(N + 1), ys(N), yp(N), yd(N);
		for (float& v : x) v = noise();
		double d0 = 0, d1 = 0;
		for (int i = 0; i < N; i++) {
			ys[i] = kb_biquad_tick(s, x[i]);
			const double y = (double)q.b0 * x[i] + d0;                       // the same filter in double
			d0 = (double)q.b1 * x[i] - (double)q.a1 * y + d1; d1 = (double)q.b2 * x[i] - (double)q.a2 * y;
			yd[i] = (float)y;
		}
		KbRv3ScanCoef sc; kb_rv3_scan_coef(q, sc);
		float z0 = 0, z1 = 0;
		for (int o = 0; o < N; o += 150) { const int ticks = std::min(150, N - o); scan_chunk_host(sc, &x[o], 0.f, ticks, &yp[o], z0, z1); }
		const double e_seq = bar_excess(ys, yd), e_scan = bar_excess(yp, ys);
		const bool adm = kb_rv3_scan_admissible(q);
		printf("%10.0f %8.4f %10s %26.4f %26.4f\n", f, q.a2, adm ? "yes" : "no", e_seq, e_scan);
		if (adm && e_scan > 0.5) fails++;
	}
	// ---------------------------------------------------------------------------------------------------- part 2
	printf("part 2: Reverb.k, 64 blocks x 4096 frames of uniform noise in [-0.5, 0.5), fs 48 kHz; worst block, in units of the bar 1e-5 |r| + 1e-6 peak\n");
	printf("%-46s %8s %22s %22s\n", "controls", "chunk", "chunked exact == A", "chunked scan vs A [bar]");
	struct Case { const char* name; float ctl[10]; } cases[] = {
		{ "default patch (early bus only)",          { 0.0f, 1.0f, 0.0f, 0.0f, 1.0f, 10.f, 1.0f, 1.0f, 1.0f, 0.f } },
		{ "all buses, damping 10 kHz / 10 kHz",      { 0.3f, 0.9f, 0.4f, 0.5f, 0.8f, 10.f, 1.0f, 1.0f, 1.0f, 0.f } },
		{ "mid + late only, damping 10 kHz / 10 kHz", { 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 10.f, 1.0f, 1.0f, 1.0f, 0.f } },
		{ "Large Hall preset (5 kHz / 2.5 kHz)",     { 1.0f, 0.0f, 0.419f, 0.329f, 1.0f, 10.f, 1.0f, 0.5f, 0.5f, 0.1f } },
		{ "mid + late only, damping 5 kHz / 5 kHz",   { 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 10.f, 0.6f, 0.5f, 1.0f, 0.f } },
		{ "mid + late only, damping 4 kHz / 4 kHz",   { 0.0f, 0.0f, 1.0f, 1.0f, 1.0f, 10.f, 0.3f, 0.4f, 1.0f, 0.f } },
	};
	FILE* dump = argc > 1 ? fopen(argv[1], "wb") : nullptr;                  // input and A of case 1 (first 8 blocks: [l, r, out l, out r] x 4096), for the comparison with the compiled reference
	for (size_t ci = 0; ci < sizeof(cases) / sizeof(cases[0]); ci++) {
		KbFxHdr hA; KbReverb A;
		memset(&hA, 0, sizeof(hA));
		kb_reverb_construct(hA, A, 0);
		for (int k = 0; k < 10; k++) kb_control_set(hA.controls[k], cases[ci].ctl[k]);
		kb_reverb_prepare(fs, hA, A);
		KbFxHdr hB = hA, hC = hA; KbReverb B = A, C = A;
		std::vector<float> ra(KB_REVERB_RING_FLOATS, 0.f), rb(KB_REVERB_RING_FLOATS, 0.f), rc(KB_REVERB_RING_FLOATS, 0.f);
		g_seed = 0x1234567ull + ci;
		bool exact_same = true, admitted = true;
		for (int k = 0; k < 2; k++) for (int i = 0; i < 4; i++) admitted = admitted && kb_rv3_scan_admissible(A.mid[k].d[i].filter) && kb_rv3_scan_admissible(A.late[k].d[i].filter);
		double worst = 0;
		const int n = 4096, chunk = plan_chunk(A);
		for (int blk = 0; blk < 64; blk++) {
			std::vector<float> l(n), r(n);
			for (int i = 0; i < n; i++) { l[i] = noise(); r[i] = noise(); }
			std::vector<float> la = l, rra = r, lb = l, rrb = r, lc = l, rrc = r;
			for (int i = 0; i < n; i++) { float ol, orr; kb_reverb_frame(hA, A, ra.data(), la[i], rra[i], ol, orr); la[i] = ol; rra[i] = orr; }
			reverb_chunked(hB, B, rb.data(), lb.data(), rrb.data(), n, false);
			reverb_chunked(hC, C, rc.data(), lc.data(), rrc.data(), n, true);
			exact_same = exact_same && memcmp(la.data(), lb.data(), n * 4) == 0 && memcmp(rra.data(), rrb.data(), n * 4) == 0;
			worst = std::max(worst, std::max(bar_excess(lc, la), bar_excess(rrc, rra)));
			if (dump && ci == 1 && blk < 8) { fwrite(l.data(), 4, n, dump); fwrite(r.data(), 4, n, dump); fwrite(la.data(), 4, n, dump); fwrite(rra.data(), 4, n, dump); }
		}
		exact_same = exact_same && memcmp(ra.data(), rb.data(), ra.size() * 4) == 0;
		printf("%-46s %8d %22s %22.4f%s\n", cases[ci].name, chunk, exact_same ? "bit-identical" : "DIFFERENT", worst, admitted ? "" : "   (not admitted: lines run the sequential filter)");
		if (!exact_same || (admitted && worst > 1.0)) fails++;
	}
	if (dump) fclose(dump);
	printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
	return fails ? 1 : 0;
}
