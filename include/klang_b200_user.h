/* klang-b200 — C ABI of ONE translated `.k` program: an Effect or a mono Synth (Tier B).
 *
 * `python -m klang_b200.kcc program.k -o libprogram_k.so` translates a klang Effect program (klang.h:4203-4217, 4703-4716: a struct derived from
 * Effect / Stereo::Effect with a controls table and a per-sample process() body), compiles ITS OWN process() body for sm_100a and links it
 * behind the entry points below.  Unlike the KB_FX_* graph ids of klang_b200.h — hand-written restatements of fixed reference programs — the
 * DSP that runs here is whatever the `.k` file says.  Conventions are those of klang_b200.h: caller-owned planar float32 buffers
 * [instances][channels][n], in place, block-granular control changes, 0 / negative return codes, no CPU implementation.
 */
#ifndef KLANG_B200_USER_H
#define KLANG_B200_USER_H
#ifdef __cplusplus
extern "C" {
#endif
const char* kb_user_name(void);                 /* the plugin struct's name */
int kb_user_channels(void);                     /* 1 = Effect, 2 = Stereo::Effect */
int kb_user_num_controls(void);                 /* size of the controls table the constructor fills   klang.h:1893-1925 */
int kb_user_stateless(void);                    /* 1: no data members -> thread-per-sample streaming kernel; 0: lane per instance, frame by frame */
const char* kb_user_last_error(void);
void* kb_user_fx_create(int instances, float fs, int max_block, int device);
void kb_user_fx_destroy(void* bank);
int kb_user_fx_set_control(void* bank, int instance, int idx, float value);      /* controls[idx].set(value): clamps   klang.h:1725-1728 */
int kb_user_fx_get_control(void* bank, int instance, int idx, float* value);
int kb_user_fx_process(void* bank, float* io, int n, unsigned flags);            /* Effect::process(buffer); flags: KB_DEVICE_PTR (1) */

/* A translated mono Synth program (klang.h:4376-4467; kb_user_kind() == 1 — the kb_user_fx_* entry points above belong to kind 0, an Effect).
 * One voice per note the program's constructor adds (notes.add<NOTE>(n), klang.h:4325-4331).  Note::on() / off() run on the host mirror
 * (NoteBase::start / release, klang.h:4257-4275), Note::process() per sample on the device, one lane per voice; process() returns
 * Synth::process(float*, int): out = [instances][n], or the per-voice streams [instances * voices][n] with flag 2 (KB_PER_VOICE). */
int kb_user_kind(void);
int kb_user_synth_voices(void);
void* kb_user_synth_create(int instances, float fs, int max_block, int device);
void kb_user_synth_destroy(void* bank);
int kb_user_synth_set_control(void* bank, int instance, int idx, float value);
int kb_user_synth_get_control(void* bank, int instance, int idx, float* value);
int kb_user_synth_note_on(void* bank, int instance, int pitch, float velocity);    /* Synth::noteOn: Notes::assign + start; returns the voice   klang.h:4423-4427, 4336-4372 */
int kb_user_synth_note_off(void* bank, int instance, int pitch, float velocity);   /* Synth::noteOff   klang.h:4430-4434 */
int kb_user_synth_voice_start(void* bank, int instance, int voice, float pitch, float velocity);
int kb_user_synth_voice_release(void* bank, int instance, int voice, float velocity);
int kb_user_synth_voice_stage(void* bank, int instance, int voice);                /* NoteBase::stage: 0 Onset, 1 Sustain, 2 Release, 3 Off */
int kb_user_synth_process(void* bank, float* out, int n, unsigned flags);
#ifdef __cplusplus
}
#endif
#endif
