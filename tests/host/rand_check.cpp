// Host-side check (g++, no CUDA) of klang_b200/csrc/kb_rand.h against the box's libc: the restated glibc TYPE_3 generator equals
// srand()/rand() draw for draw, the jump-ahead equals sequential stepping, and capture / commit attach to the LIVE libc stream
// (rand() continues exactly where the advanced state says, and untouched after a bare capture).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../klang_b200/csrc/kb_rand.h"

int main() {
	long long bad = 0, draws = 0;
	const unsigned seeds[] = { 0u, 1u, 2u, 7u, 272839u, 12345u, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu, 2147483647u / 16807u, 3141592653u };
	for (unsigned seed : seeds) {                                           // 1. generator and seeding
		srand(seed);
		KbRand g; kb_rand_seed(g, seed);
		for (int i = 0; i < 200000; i++) { draws++; if ((int)kb_rand_next(g) != rand()) bad++; }
	}
	long long jump_bad = 0;
	{                                                                       // 2. jump == sequential, state for state
		const unsigned long long ns[] = { 0, 1, 30, 31, 32, 123, 124, 125, 4096, 65537, 1000003, 4194304, 123456789ull };
		for (unsigned seed = 1; seed <= 3; seed++)
			for (unsigned long long n : ns) {
				KbRand a, b; kb_rand_seed(a, seed * 977u); for (unsigned i = 0; i < seed * 13; i++) kb_rand_next(a);
				b = a;
				for (unsigned long long i = 0; i < n; i++) kb_rand_next(a);
				kb_rand_jump(b, n);
				for (int i = 0; i < 100; i++) if (kb_rand_next(a) != kb_rand_next(b)) jump_bad++;
			}
	}
	long long live_bad = 0;
	{                                                                       // 3. the live libc stream
		srand(99); for (int i = 0; i < 1234; i++) rand();
		KbRand g;
		if (!kb_rand_capture(g)) live_bad += 1000000;
		KbRand peek = g;
		for (int i = 0; i < 500; i++) if ((int)kb_rand_next(peek) != rand()) live_bad++;       // capture left libc untouched and saw its position
		if (!kb_rand_capture(g)) live_bad += 1000000;
		// a "device block" consumes 3 x 4096 draws: the host jumps its copy, commits, and libc rand() continues after them
		KbRand ref = g; for (int i = 0; i < 3 * 4096; i++) kb_rand_next(ref);
		kb_rand_jump(g, 3 * 4096);
		if (!kb_rand_commit(g)) live_bad += 1000000;
		for (int i = 0; i < 5000; i++) if ((int)kb_rand_next(ref) != rand()) live_bad++;
		srand(5);                                                           // srand() after a commit still reseeds the same (default) stream
		KbRand s5; kb_rand_seed(s5, 5);
		for (int i = 0; i < 1000; i++) if ((int)kb_rand_next(s5) != rand()) live_bad++;
	}
	// 4. the noise maps (klang.h:4947-4951, 5357-5366) against the expressions as the reference writes them
	long long noise_bad = 0;
	srand(3);
	for (int i = 0; i < 100000; i++) {
		const int r = rand();
		const float basic = r * 2.f / (const float)RAND_MAX - 1.f;
		union { unsigned int u; float f; } w; w.u = ((r & 0b111111111111111UL) << 1) | 0b1000011100000000000000000000000u;
		const float fast = w.f - 257.f;
		const float b2 = kb_noise_basic((uint32_t)r), f2 = kb_noise_fast((uint32_t)r);
		if (memcmp(&basic, &b2, 4) || memcmp(&fast, &f2, 4)) noise_bad++;
	}
	printf("rand check: %lld draws, %lld mismatches, %lld jump mismatches, %lld live-stream mismatches, %lld noise mismatches\n", draws, bad, jump_bad, live_bad, noise_bad);
	return (bad || jump_bad || live_bad || noise_bad) ? 1 : 0;
}
