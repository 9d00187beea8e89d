/* klang-b200 — C ABI of ONE translated `.k` Effect (Tier B).
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
#ifdef __cplusplus
}
#endif
#endif
