// klang-b200 — libc rand() on the device.
//
// The reference draws libc rand() from per-sample code (Generators::Basic::Noise klang.h:4947-4951, Fast::Noise 5357-5366) and
// from event code (klang::random, klang.h:237-240), all from the ONE process-wide stream, so a device kernel that produces
// noise has to continue that stream bit for bit and hand it back advanced (SURVEY Q9, §8f-1 "a device rand() bit-matching
// glibc").  glibc 2.39 rand() = random() with the default TYPE_3 state (stdlib/random.c, random_r.c — third-party, not under
// /root/reference; restated from its published algorithm and pinned against the box's libc by tests/host/rand_check.cpp):
// an additive lagged-Fibonacci generator over 31 words, y[k] = y[k-3] + y[k-31] (mod 2^32), output y[k] >> 1; srand(seed)
// fills the words with the Lehmer generator 16807 (Schrage form, int32) and discards 310 draws.
//
//   kb_rand_seed / kb_rand_next      the generator itself (host + device)
//   kb_rand_jump                     advance by n draws in O(31^2 log n): the recurrence is linear over Z/2^32, so
//                                    x^n mod (x^31 - x^28 - 1) gives the n-draws-later state — how a kernel hands every voice its
//                                    own slice of the stream and how the host catches up with what the device consumed
//   kb_rand_capture / kb_rand_commit (host) read the process's live libc state and write it back advanced, through the documented
//                                    initstate()/setstate() hand-over: both stamp the outgoing buffer's first word with its rear
//                                    index and setstate() reads the index back from there (random_r.c)
#pragma once
#include <cstdint>
#include <cstdlib>
#ifndef KB_HD
#ifdef __CUDACC__
#define KB_HD __host__ __device__ inline
#else
#define KB_HD inline
#endif
#endif

#define KB_RAND_DEG 31
#define KB_RAND_SEP 3

struct KbRand { uint32_t s[KB_RAND_DEG]; int rear; };      // glibc's state[0..30] and (rptr - state); fptr = rear + 3 (mod 31)

KB_HD uint32_t kb_rand_next(KbRand& g) {                   // random_r(): *fptr += *rptr; result = *fptr >> 1; advance both
	int f = g.rear + KB_RAND_SEP; if (f >= KB_RAND_DEG) f -= KB_RAND_DEG;
	const uint32_t v = (g.s[f] += g.s[g.rear]);
	g.rear = (g.rear + 1 == KB_RAND_DEG) ? 0 : g.rear + 1;
	return v >> 1;
}
KB_HD void kb_rand_seed(KbRand& g, unsigned seed) {        // srandom_r()
	if (seed == 0) seed = 1;
	int32_t word = (int32_t)seed;
	g.s[0] = (uint32_t)word;
	for (int i = 1; i < KB_RAND_DEG; i++) {
		const int32_t hi = word / 127773, lo = word % 127773;
		word = 16807 * lo - 2836 * hi;
		if (word < 0) word += 2147483647;
		g.s[i] = (uint32_t)word;
	}
	g.rear = 0;
	for (int i = 0; i < 10 * KB_RAND_DEG; i++) kb_rand_next(g);
}
// Generators::Basic::Noise::process / Fast::Noise::process on one draw r = rand()
KB_HD float kb_noise_basic(uint32_t r) { return (float)(int)r * 2.f / 2147483648.0f - 1.f; }          // (const float)RAND_MAX rounds to 2^31
KB_HD float kb_noise_fast(uint32_t r) {
	const uint32_t i = ((r & 0x7fffu) << 1) | 0x43800000u;                                              // bias 0b1000011100000000000000000000000
#ifdef __CUDA_ARCH__
	return __uint_as_float(i) - 257.f;
#else
	float f; __builtin_memcpy(&f, &i, 4); return f - 257.f;
#endif
}

// ---- jump-ahead (host).  Y[t] = y[k-31+t], t = 0..30, are the 31 latest values in age order; slot of Y[t] = (rear + 3 + t) mod 31.
inline void kb_rand_poly_mulmod(const uint32_t* a, const uint32_t* b, uint32_t* out) {   // out = a * b mod (x^31 - x^28 - 1) over Z/2^32
	uint32_t p[2 * KB_RAND_DEG - 1] = { 0 };
	for (int i = 0; i < KB_RAND_DEG; i++) if (a[i]) for (int j = 0; j < KB_RAND_DEG; j++) p[i + j] += a[i] * b[j];
	for (int d = 2 * KB_RAND_DEG - 2; d >= KB_RAND_DEG; d--) { p[d - KB_RAND_SEP] += p[d]; p[d - KB_RAND_DEG] += p[d]; }   // x^d = x^(d-3) + x^(d-31)
	for (int i = 0; i < KB_RAND_DEG; i++) out[i] = p[i];
}
inline void kb_rand_jump(KbRand& g, unsigned long long n) {
	if (n < 4 * KB_RAND_DEG) { for (unsigned long long i = 0; i < n; i++) kb_rand_next(g); return; }
	uint32_t c[KB_RAND_DEG] = { 1 }, x[KB_RAND_DEG] = { 0, 1 }, t[KB_RAND_DEG];              // c = x^n mod P by square and multiply
	for (unsigned long long e = n; e; e >>= 1) {
		if (e & 1) { kb_rand_poly_mulmod(c, x, t); for (int i = 0; i < KB_RAND_DEG; i++) c[i] = t[i]; }
		kb_rand_poly_mulmod(x, x, t); for (int i = 0; i < KB_RAND_DEG; i++) x[i] = t[i];
	}
	uint32_t y[2 * KB_RAND_DEG - 1];                                                         // y[k-31 .. k+29]: the state and the next 30 values
	for (int i = 0; i < KB_RAND_DEG; i++) y[i] = g.s[(g.rear + KB_RAND_SEP + i) % KB_RAND_DEG];
	for (int i = KB_RAND_DEG; i < 2 * KB_RAND_DEG - 1; i++) y[i] = y[i - KB_RAND_SEP] + y[i - KB_RAND_DEG];
	const int rear = (int)((g.rear + n) % KB_RAND_DEG);
	for (int i = 0; i < KB_RAND_DEG; i++) {                                                  // y[k+n-31+i] = sum_m c[m] y[k-31+i+m]
		uint32_t acc = 0;
		for (int m = 0; m < KB_RAND_DEG; m++) acc += c[m] * y[i + m];
		g.s[(rear + KB_RAND_SEP + i) % KB_RAND_DEG] = acc;
	}
	g.rear = rear;
}

// ---- the process's libc stream (host).  Returns false when libc is not running the default TYPE_3 generator (an application that
// installed a state of another size with initstate()): the caller then reports "unsupported" instead of guessing.
// Buffer layout of initstate()/setstate() (random.c, random_r.c): word 0 = 5 * rear + type, words 1..31 = the state; both calls return
// the outgoing buffer and stamp its word 0 first.
inline bool kb_rand_capture(KbRand& g) {
	alignas(8) static char scratch[128];
	int32_t* live = reinterpret_cast<int32_t*>(initstate(1u, scratch, sizeof(scratch)));     // switches libc to `scratch`
	if (!live) return false;
	const int32_t tag = live[0];
	const bool ok = tag % 5 == 3 && tag / 5 >= 0 && tag / 5 < KB_RAND_DEG;
	if (ok) { for (int i = 0; i < KB_RAND_DEG; i++) g.s[i] = (uint32_t)live[1 + i]; g.rear = tag / 5; }
	setstate(reinterpret_cast<char*>(live));                                                 // back to the live state, untouched
	return ok;
}
inline bool kb_rand_commit(const KbRand& g) {
	alignas(8) static char scratch[128];
	int32_t* live = reinterpret_cast<int32_t*>(initstate(1u, scratch, sizeof(scratch)));
	if (!live) return false;
	const bool ok = live[0] % 5 == 3;
	if (ok) { for (int i = 0; i < KB_RAND_DEG; i++) live[1 + i] = (int32_t)g.s[i]; live[0] = 5 * g.rear + 3; }
	setstate(reinterpret_cast<char*>(live));                                                 // reads the type and the rear index back from word 0
	return ok;
}
