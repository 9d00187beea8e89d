// Host-side check (g++, no CUDA): the product's Delay<1000> known-answer loop (kb_delay_kat of klang_b200/csrc/kb_prims.cuh — the
// function kb_prim_delay_kernel runs) on the inputs found in argv[1] (raw: int32 n, then in[n], di[n] (int32), df[n], set_at[n]);
// writes out_i, out_f, out_p, out_l as raw float32 to stdout.  tests/test_host_logic.py compares them with the golden vectors of
// the compiled reference.
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
	int32_t n = 0;
	if (fread(&n, 4, 1, f) != 1 || n <= 0) return 2;
	std::vector<float> in(n), df(n), set_at(n), ring(1001), oi(n), of(n), op(n), ol(n);
	std::vector<int> di(n);
	if (fread(in.data(), 4, n, f) != (size_t)n || fread(di.data(), 4, n, f) != (size_t)n || fread(df.data(), 4, n, f) != (size_t)n ||
	    fread(set_at.data(), 4, n, f) != (size_t)n) return 2;
	fclose(f);
	kb_delay_kat(n, in.data(), di.data(), df.data(), set_at.data(), ring.data(), oi.data(), of.data(), op.data(), ol.data());
	fwrite(oi.data(), 4, n, stdout); fwrite(of.data(), 4, n, stdout); fwrite(op.data(), 4, n, stdout); fwrite(ol.data(), 4, n, stdout);
	return 0;
}
