/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Plain-C restatement of the klang v0.7.8 per-sample signal-flow path
 * (reference: nashaudio/klang, klang.h + examples/NAME.k).  Every function cites the
 * reference file:line it follows.  The restatement is validated bit-for-bit against
 * the compiled reference (oracle/_ref/libklang_ref.so) by tests/test_oracle_port.py
 * and against the committed golden vectors under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * Like the reference it keeps process-global state (sample rate, libc rand()).
 */
#ifndef KLANG_PORT_H
#define KLANG_PORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ globals */
void  kp_set_fs(float fs);              /* klang::fs = SampleRate(fs)   klang.h:1593-1604 */
float kp_get_fs(void);
void  kp_srand(unsigned seed);          /* klang::random(seed)          klang.h:240 */
int   kp_version(void);
float kp_pitch_to_frequency(float p);   /* Pitch::operator->            klang.h:1568-1571 */

/* --------------------------------------------------------------- primitives */
int kp_osc(int kind, int nargs, float f, float phase, float duty, int n, float* out);
int kp_wavetable(int kind, float* table);
int kp_filter(int kind, int nset, const float* f, const float* Q, int n, const float* in, float* out, float* coeffs);
int kp_envelope(int npts, const float* xy, int loop_start, int loop_end, int n, int release_at,
                float release_time, float release_level, float* out, int* stage_out);
int kp_envelope_at(int npts, const float* xy, int n, const float* t, float* out);
int kp_adsr(float A, float D, float S, float R, int n, int release_at, float* out, int* stage_out);
int kp_delay1000(int n, const float* in, const int* di, const float* df, const float* set_at,
                 float* out_i, float* out_f, float* out_p);
int kp_stereo_delay1000(int n, const float* inl, const float* inr, const float* df, float* outl, float* outr);
int kp_control_smooth(float lo, float hi, float initial, int n, const float* values, float* out);

int kp_sample(const float* table, int size, int nargs, float f, float phase, int n, float* out);       /* klang::Sample  klang.h:3679-3720 */
int kp_wav_decode(const unsigned char* bytes, long long nbytes, float* out, int max, int* info);     /* File::WAV      klang.h:5997-6085 */

/* ------------------------------------------------------------------ effects */
void* kp_fx_create(int graph);
void  kp_fx_destroy(void* h);
int   kp_fx_channels(void* h);
int   kp_fx_num_controls(void* h);
void  kp_fx_set_control(void* h, int idx, float v);
float kp_fx_get_control(void* h, int idx);
int   kp_fx_process(void* h, float* l, float* r, int n);
int   kp_fx_debug(void* h, float* dst, int n);   /* the block's `>> debug` capture (klang.h:3132-3287); 1 if written */

int   kp_fx_num_presets(void* h);           /* Plugin::presets  klang.h:1940-1981, 4195-4200 */
int   kp_fx_preset(void* h, int p, char* name, int name_max, float* values, int max);
int   kp_fx_load_preset(void* h, int p);    /* values through Control::set, then onPreset  klang.h:4190 */

/* ------------------------------------------------------------------- synths */
void* kp_synth_create(int graph, int nvoices);
void  kp_synth_destroy(void* h);
int   kp_synth_channels(void* h);
int   kp_synth_num_voices(void* h);
int   kp_synth_num_controls(void* h);
void  kp_synth_set_control(void* h, int idx, float v);
float kp_synth_get_control(void* h, int idx);
int   kp_synth_note_on(void* h, int pitch, float velocity);
void  kp_synth_note_off(void* h, int pitch, float velocity);
void  kp_synth_voice_start(void* h, int voice, float pitch, float velocity);
void  kp_synth_voice_release(void* h, int voice, float velocity);
int   kp_synth_voice_stage(void* h, int voice);
int   kp_synth_num_presets(void* h);
int   kp_synth_preset(void* h, int p, char* name, int name_max, float* values, int max);
int   kp_synth_load_preset(void* h, int p);
int   kp_synth_on_control(void* h, int idx, float value);   /* Synth::onControl  klang.h:4399-4404; returns the notes notified */
int   kp_synth_process(void* h, float* l, float* r, int n);
int   kp_synth_process_voices(void* h, float* out, int n, int* active);

#ifdef __cplusplus
}
#endif
#endif
