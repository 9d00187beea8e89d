// Host program: the reference's example `.k` programs, UNMODIFIED, compiled against include/compat/klang.h and run on the
// B200 through libklang_b200.so.  Built by tools/build_k_host.py from the sources where they lie under
// /root/reference/examples into tests/_k_bin/ (git-ignored build artefact, like oracle/_ref); used by
// tests/test_k_programs.py.  Usage: k_host <program> <fs> <block> <blocks> <out.f32>
#include <klang.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace k_gain {
#include "Gain/Gain.k"
}
namespace k_pingpong {
#include "PingPong.k"
}
namespace k_delay_pingpong {
#include "Delay/PingPong.k"
}
namespace k_delay_reverb {
#include "Delay/Reverb.k"
}
namespace k_supersaw {
#include "SuperSaw.k"
}
namespace k_filter {
#include "Subtractive/Filter.k"
}
namespace k_tb303 {
#include "TB303.k"
}
namespace k_reverb {
#include "Reverb.k"
}
namespace k_synthx {
#include "SynTHX.k"
}
namespace k_fm {
#include "FM.k"
}
namespace k_pan {
#include "Gain/Pan.k"
}
namespace k_rm {
#include "Gain/RM.k"
}
namespace k_tremolo {
#include "Gain/Tremolo.k"
}
namespace k_clipping {
#include "Distortion/Clipping.k"
}
namespace k_echo {
#include "Delay/Echo.k"
}
namespace k_feedback {
#include "Delay/Feedback.k"
}
namespace k_add_saw {
#include "Additive/Saw.k"
}
namespace k_add_square {
#include "Additive/Square.k"
}
namespace k_am {
#include "Modulation/AM.k"
}
namespace k_mod_fm {
#include "Modulation/FM.k"
}
namespace k_mod_fm2 {
#include "Modulation/FM2.k"
}
namespace k_functions {
#include "Distortion/Functions.k"
}
namespace k_mute {
#include "Distortion/Mute.k"
}
namespace k_add_nyquist {
#include "Additive/Nyquist.k"
}
namespace k_iir {
#include "Filtering/IIR.k"
}
namespace k_wahwah {
#include "Filtering/WahWah.k"
}
namespace k_flanger {
#include "Modulation/Flanger.k"
}
namespace k_moddelay {
#include "Modulation/ModDelay.k"
}
namespace k_mod_chorus {
#include "Modulation/Chorus.k"
}
namespace k_breakpoint {
#include "Subtractive/Breakpoint.k"
}
namespace k_ramp {
#include "Subtractive/Ramp.k"
}
namespace k_release {
#include "Subtractive/Release.k"
}

KLANG_B200_EFFECT(k_gain::Gain, KB_FX_GAIN)
KLANG_B200_EFFECT(k_pingpong::PingPong, KB_FX_PINGPONG)
KLANG_B200_EFFECT(k_delay_pingpong::PingPong, KB_FX_DELAY_PINGPONG)
KLANG_B200_EFFECT(k_delay_reverb::Reverb, KB_FX_DELAY_REVERB)
KLANG_B200_SYNTH(k_supersaw::SuperSaw, KB_SY_SUPERSAW)
KLANG_B200_SYNTH(k_filter::Filter, KB_SY_FILTER_K)
KLANG_B200_SYNTH(k_tb303::TB303, KB_SY_TB303)
KLANG_B200_EFFECT(k_reverb::Reverb, KB_FX_REVERB)
KLANG_B200_SYNTH(k_synthx::SynTHX, KB_SY_SYNTHX)
KLANG_B200_SYNTH(k_fm::FM, KB_SY_FM)
KLANG_B200_EFFECT(k_pan::Pan, KB_FX_PAN)
KLANG_B200_EFFECT(k_rm::RM, KB_FX_RM)
KLANG_B200_EFFECT(k_tremolo::Tremolo, KB_FX_TREMOLO)
KLANG_B200_EFFECT(k_clipping::Clipping, KB_FX_CLIPPING)
KLANG_B200_EFFECT(k_echo::Echo, KB_FX_ECHO)
KLANG_B200_EFFECT(k_feedback::Feedback, KB_FX_FEEDBACK)
KLANG_B200_SYNTH(k_add_saw::Saw, KB_SY_ADDITIVE_SAW)
KLANG_B200_SYNTH(k_add_square::Square, KB_SY_ADDITIVE_SQUARE)
KLANG_B200_SYNTH(k_am::AM, KB_SY_AM)
KLANG_B200_SYNTH(k_mod_fm::FM, KB_SY_MOD_FM)
KLANG_B200_SYNTH(k_mod_fm2::FM2, KB_SY_MOD_FM2)
KLANG_B200_EFFECT(k_functions::Functions, KB_FX_FUNCTIONS)
KLANG_B200_EFFECT(k_mute::Mute, KB_FX_MUTE)
KLANG_B200_SYNTH(k_add_nyquist::Nyquist, KB_SY_ADDITIVE_NYQUIST)
KLANG_B200_EFFECT(k_iir::IIR, KB_FX_IIR)
KLANG_B200_EFFECT(k_wahwah::WahWah, KB_FX_WAHWAH)
KLANG_B200_EFFECT(k_flanger::Flanger, KB_FX_FLANGER)
KLANG_B200_EFFECT(k_moddelay::ModDelay, KB_FX_MODDELAY)
KLANG_B200_EFFECT(k_mod_chorus::Chorus, KB_FX_MOD_CHORUS)
KLANG_B200_SYNTH(k_breakpoint::Breakpoint, KB_SY_BREAKPOINT)
KLANG_B200_SYNTH(k_ramp::Ramp, KB_SY_RAMP)
KLANG_B200_SYNTH(k_release::Release, KB_SY_RELEASE)

// the deterministic input of tests/cases.py::noise
static float noise(uint64_t n, uint64_t seed, double lo, double hi) {
	uint64_t x = (n + seed * 0x9E3779B97F4A7C15ull) * 6364136223846793005ull + 1442695040888963407ull;
	x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33;
	return (float)(lo + (hi - lo) * ((double)(x >> 40) / 16777216.0));
}

template <class PLUGIN> static int run_effect(float fs, int n, int blocks, FILE* out) {
	klang::b200::EffectHost<PLUGIN> host(fs, n);
	const int C = host.channels();
	std::vector<float> buf((size_t)C * n);
	for (int b = 0; b < blocks; b++) {
		if (b == 1) host.plugin.controls[0].set(0.3f);            // the host moves a control between blocks, through the .k object
		for (int c = 0; c < C; c++) for (int t = 0; t < n; t++) buf[(size_t)c * n + t] = noise((uint64_t)b * n + t, 2 + c, -0.5, 0.5);
		if (!host.process(buf.data(), n)) return 2;
		fwrite(buf.data(), sizeof(float), buf.size(), out);
	}
	return 0;
}
template <class PLUGIN> static int run_synth(float fs, int n, int blocks, FILE* out) {
	klang::b200::SynthHost<PLUGIN> host(fs, n);
	kb_srand(1);
	const int C = host.channels();
	std::vector<float> buf((size_t)C * n);
	for (int b = 0; b < blocks; b++) {
		if (b == 0) for (int k = 0; k < 6; k++) host.noteOn(48 + 5 * k, 0.5f + 0.08f * k);
		if (b == 2) { host.noteOff(48); host.noteOff(58); host.noteOn(77, 0.9f); }
		if (!host.process(buf.data(), n)) return 2;
		fwrite(buf.data(), sizeof(float), buf.size(), out);
	}
	return 0;
}

int main(int argc, char** argv) {
	if (argc < 6) { fprintf(stderr, "usage: k_host <program> <fs> <block> <blocks> <out.f32>\n"); return 64; }
	const std::string prog = argv[1];
	const float fs = (float)atof(argv[2]);
	const int n = atoi(argv[3]), blocks = atoi(argv[4]);
	FILE* out = fopen(argv[5], "wb");
	if (!out) return 65;
	int rc = 66;
	try {
		if (prog == "gain") rc = run_effect<k_gain::Gain>(fs, n, blocks, out);
		else if (prog == "pingpong") rc = run_effect<k_pingpong::PingPong>(fs, n, blocks, out);
		else if (prog == "delay_pingpong") rc = run_effect<k_delay_pingpong::PingPong>(fs, n, blocks, out);
		else if (prog == "delay_reverb") rc = run_effect<k_delay_reverb::Reverb>(fs, n, blocks, out);
		else if (prog == "supersaw") rc = run_synth<k_supersaw::SuperSaw>(fs, n, blocks, out);
		else if (prog == "filter_k") rc = run_synth<k_filter::Filter>(fs, n, blocks, out);
		else if (prog == "tb303") rc = run_synth<k_tb303::TB303>(fs, n, blocks, out);
		else if (prog == "reverb") rc = run_effect<k_reverb::Reverb>(fs, n, blocks, out);
		else if (prog == "synthx") rc = run_synth<k_synthx::SynTHX>(fs, n, blocks, out);
		else if (prog == "fm") rc = run_synth<k_fm::FM>(fs, n, blocks, out);
		else if (prog == "pan") rc = run_effect<k_pan::Pan>(fs, n, blocks, out);
		else if (prog == "rm") rc = run_effect<k_rm::RM>(fs, n, blocks, out);
		else if (prog == "tremolo") rc = run_effect<k_tremolo::Tremolo>(fs, n, blocks, out);
		else if (prog == "clipping") rc = run_effect<k_clipping::Clipping>(fs, n, blocks, out);
		else if (prog == "echo") rc = run_effect<k_echo::Echo>(fs, n, blocks, out);
		else if (prog == "feedback") rc = run_effect<k_feedback::Feedback>(fs, n, blocks, out);
		else if (prog == "additive_saw") rc = run_synth<k_add_saw::Saw>(fs, n, blocks, out);
		else if (prog == "additive_square") rc = run_synth<k_add_square::Square>(fs, n, blocks, out);
		else if (prog == "am") rc = run_synth<k_am::AM>(fs, n, blocks, out);
		else if (prog == "mod_fm") rc = run_synth<k_mod_fm::FM>(fs, n, blocks, out);
		else if (prog == "mod_fm2") rc = run_synth<k_mod_fm2::FM2>(fs, n, blocks, out);
		else if (prog == "functions") rc = run_effect<k_functions::Functions>(fs, n, blocks, out);
		else if (prog == "mute") rc = run_effect<k_mute::Mute>(fs, n, blocks, out);
		else if (prog == "additive_nyquist") rc = run_synth<k_add_nyquist::Nyquist>(fs, n, blocks, out);
		else if (prog == "iir") rc = run_effect<k_iir::IIR>(fs, n, blocks, out);
		else if (prog == "wahwah") rc = run_effect<k_wahwah::WahWah>(fs, n, blocks, out);
		else if (prog == "flanger") rc = run_effect<k_flanger::Flanger>(fs, n, blocks, out);
		else if (prog == "moddelay") rc = run_effect<k_moddelay::ModDelay>(fs, n, blocks, out);
		else if (prog == "mod_chorus") rc = run_effect<k_mod_chorus::Chorus>(fs, n, blocks, out);
		else if (prog == "breakpoint") rc = run_synth<k_breakpoint::Breakpoint>(fs, n, blocks, out);
		else if (prog == "ramp") rc = run_synth<k_ramp::Ramp>(fs, n, blocks, out);
		else if (prog == "release") rc = run_synth<k_release::Release>(fs, n, blocks, out);
	} catch (const klang::b200::Error& e) {
		fprintf(stderr, "k_host: %s\n", e.what());
		rc = kb_device_count() == 0 ? 3 : 4;          // 3 = no CUDA device (expected off the GPU box)
	} catch (const std::exception& e) {
		fprintf(stderr, "k_host: %s\n", e.what());
		rc = 5;
	}
	fclose(out);
	return rc;
}
