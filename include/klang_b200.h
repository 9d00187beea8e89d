/* klang-b200 — C ABI of the B200-native klang hot path.
 *
 * The reference (nashaudio/klang v0.7.8) has no FFI boundary: its hot path is entered through the C++
 * block drivers of klang.h.  Each entry point below names the reference interface it replaces
 * (klang.h:line); INTEGRATION.md shows the binding a klang maintainer adds to route
 * Effect::process / Synth::process through this library.
 *
 * A *bank* is many independent instances of one klang plugin (Effect or Synth) evaluated together by
 * hand-written sm_100a kernels; voice / instance state lives in HBM between calls.  Conventions follow the
 * reference: buffers are caller-owned planar float32, effects run in place, events (controls, notes) are
 * block-granular and take effect at the next process call in call order, there are no exceptions.
 * Every function that can fail returns 0 on success or a negative KB_E* code (the reference returns
 * void); kb_last_error() describes the last failure on the calling thread.  One caller thread per bank.
 * There is no CPU implementation behind these calls: without a CUDA device create() fails.
 */
#ifndef KLANG_B200_H
#define KLANG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define KB_VERSION 100            /* library 0.1.0 */
#define KB_KLANG_VERSION 708      /* restates klang v0.7.8 (klang.h:85) */

/* error codes */
#define KB_OK 0
#define KB_EINVAL (-1)            /* bad argument (range, null, block > max_block) */
#define KB_ECUDA (-2)             /* CUDA runtime error, see kb_last_error() */
#define KB_ENODEV (-3)            /* no usable CUDA device */

/* effect graphs (the user .k program each one reproduces) */
#define KB_FX_GAIN 0              /* examples/Gain/Gain.k          mono   */
#define KB_FX_PINGPONG 1          /* examples/PingPong.k           stereo */
#define KB_FX_REVERB 2            /* examples/Reverb.k             stereo */
#define KB_FX_DELAY_PINGPONG 3    /* examples/Delay/PingPong.k     stereo */
#define KB_FX_DELAY_REVERB 4      /* examples/Delay/Reverb.k       mono   */
#define KB_FX_PAN 5               /* examples/Gain/Pan.k           stereo (elementwise) */
#define KB_FX_RM 6                /* examples/Gain/RM.k            mono   (elementwise, Fast::Sine LFO) */
#define KB_FX_TREMOLO 7           /* examples/Gain/Tremolo.k       mono   (elementwise, Fast::Sine LFO) */
#define KB_FX_CLIPPING 8          /* examples/Distortion/Clipping.k mono  (elementwise) */
#define KB_FX_ECHO 9              /* examples/Delay/Echo.k         mono   (one Delay<192000>, feed-forward tap) */
#define KB_FX_FEEDBACK 10         /* examples/Delay/Feedback.k     mono   (one Delay<192000> fed the output) */
#define KB_FX_FUNCTIONS 11        /* examples/Distortion/Functions.k mono (elementwise: hardclip(in * gain)) */
#define KB_FX_MUTE 12             /* examples/Distortion/Mute.k    mono   (elementwise: a Toggle) */
#define KB_FX_IIR 13              /* examples/Filtering/IIR.k      mono   (one-pole smoother, coefficient cube(control)) */
#define KB_FX_WAHWAH 14           /* examples/Filtering/WahWah.k   mono   (Biquad::LPF, cutoff set every sample from a sine LFO) */
#define KB_FX_FLANGER 15          /* examples/Modulation/Flanger.k  mono  (delay tapped at a triangle-LFO time, plus the input) */
#define KB_FX_MODDELAY 16         /* examples/Modulation/ModDelay.k mono  (delay tapped at a sine-LFO time, smoothed depth control) */
#define KB_FX_MOD_CHORUS 17       /* examples/Modulation/Chorus.k   mono  (one delay, three sine-LFO taps) */

/* synth graphs */
#define KB_SY_SUBTRACTIVE 0       /* Saw >> LPF(env) >> ADSR: Filter.k with a Saw and ADSR controls mono */
#define KB_SY_SUPERSAW 1          /* examples/SuperSaw.k           mono   */
#define KB_SY_TB303 2             /* examples/TB303.k              mono   */
#define KB_SY_SYNTHX 3            /* examples/SynTHX.k             stereo */
#define KB_SY_FILTER_K 4          /* examples/Subtractive/Filter.k mono   */
#define KB_SY_FM 5                /* examples/FM.k (three Operator<Sine> in series) mono */
#define KB_SY_BREAKPOINT 6        /* examples/Subtractive/Breakpoint.k (Sine x attack / decay envelope; noteOff cuts the note) mono */
#define KB_SY_RAMP 7              /* examples/Subtractive/Ramp.k       (Sine x one ramp) mono */
#define KB_SY_RELEASE 8           /* examples/Subtractive/Release.k    (Sine x looped envelope with release()) mono */
#define KB_SY_ADDITIVE_SAW 9      /* examples/Additive/Saw.k    (32 Sine partials summed in order) mono */
#define KB_SY_ADDITIVE_SQUARE 10  /* examples/Additive/Square.k (odd partials below Nyquist) mono */
#define KB_SY_AM 11               /* examples/Modulation/AM.k  (sine carrier x sine modulator, ADSR) mono */
#define KB_SY_MOD_FM 12           /* examples/Modulation/FM.k  (carrier frequency set every sample from a sine modulator) mono */
#define KB_SY_MOD_FM2 13          /* examples/Modulation/FM2.k (two modulators in series) mono */
#define KB_SY_ADDITIVE_NYQUIST 14 /* examples/Additive/Nyquist.k (every partial below Nyquist) mono */

/* process flags */
#define KB_DEVICE_PTR 1u          /* `io` / `out` is device memory on the bank's device; the call is asynchronous on the bank stream */
#define KB_PER_VOICE 2u           /* synth: write every voice alone, out = [instances][voices][channels][n] (parity / debugging) */
#define KB_MIX_SUM 4u             /* synth: mono synths sum their voices (Stereo::Note rule, klang.h:4731) instead of the reference's
                                     overwrite (klang.h:4299) */
#define KB_LANE_PER_VOICE 16u     /* synth: use the plain lane-per-voice schedule instead of the tiled one (A/B measurement; same results) */
#define KB_FX_SEQUENTIAL 32u       /* effects: use only the frame-sequential schedule (A/B measurement; same results) */
#define KB_FX_TOLERANCE 128u      /* effects: allow schedules that RE-ASSOCIATE fp32 recurrences (parallel scans over IIR filters) where the result provably
                                     stays within the parity bar 1e-5 |r| + 1e-6 peak of the reference (BASELINE north_star).  Default off: every
                                     schedule is bit-identical to the reference.  Today: the 16 damping low-passes of Reverb.k (kb_reverb3.cuh) */
#define KB_BANK_MIX 8u            /* synth: additionally sum all instances, out = [channels][n] (the multi-GPU mix-down input) */
#define KB_ASYNC_HOST 64u         /* host-pointer call: return once the work and the device-to-host copy are queued on the bank stream;
                                     `io` / `out` must be page-locked and is valid after kb_*_bank_sync (or an event the caller
                                     records on its stream).  Lets a host prepare block k+1's events while block k renders.
                                     (A page-locked, device-mapped `out` of a synth bank's KB_BANK_MIX | KB_MIX_SUM call is written by the mix
                                     kernel itself — no copy is queued — with or without this flag; the bytes and their validity are the same.) */

typedef struct kb_fx_bank kb_fx_bank;
typedef struct kb_synth_bank kb_synth_bank;

/* ------------------------------------------------------------------------------------------- library */
int kb_version(void);
int kb_device_count(void);                          /* 0 when no CUDA device / driver is usable */
const char* kb_last_error(void);                    /* thread-local, never NULL */
void kb_srand(unsigned seed);                       /* klang::random(seed) = srand()            klang.h:240 */
float kb_pitch_to_frequency(float pitch);           /* Pitch::operator-> Frequency              klang.h:1568-1571 */
/* The klang program a graph id restates: path under the reference tree and the hash of its source (comments removed, white space collapsed,
 * FNV-1a 64 — tools/k_hash.py).  A host that compiled a `.k` file binds it to the id only if the hashes agree (include/compat/klang.h):
 * an edited program must never silently run the unedited DSP.  synth = 0: KB_FX_* ids, 1: KB_SY_* ids; hash 0 = no single program. */
unsigned long long kb_graph_source_hash(int synth, int graph);
const char* kb_graph_source_path(int synth, int graph);

/* ------------------------------------------------------------------------------------ effect banks */
/* `instances` objects of the Effect / Stereo::Effect subclass `graph`, constructed with klang::fs = fs
 * (klang.h:1593-1604).  max_block bounds `n` of process().  device = CUDA ordinal. */
kb_fx_bank* kb_fx_bank_create(int graph, int instances, float fs, int max_block, int device);
void kb_fx_bank_destroy(kb_fx_bank* bank);
int kb_fx_bank_channels(const kb_fx_bank* bank);
int kb_fx_bank_instances(const kb_fx_bank* bank);
int kb_fx_bank_num_controls(const kb_fx_bank* bank);
/* controls[idx].set(value) — clamps to the control's range                               klang.h:1725-1728, 4444-4447 */
int kb_fx_bank_set_control(kb_fx_bank* bank, int instance, int idx, float value);
/* params[c] = controls[c].value write-back (PingPong.k modifies a control per sample)    klang.h:4462-4465 */
int kb_fx_bank_get_control(kb_fx_bank* bank, int instance, int idx, float* value);
/* Effect::process(buffer) / Stereo::Effect::process(Stereo::buffer) for every instance, in place.
 * io = [instances][channels][n] planar float32.                                          klang.h:4208-4216, 4708-4716 */
int kb_fx_bank_process(kb_fx_bank* bank, float* io, int n, unsigned flags);
int kb_fx_bank_sync(kb_fx_bank* bank);              /* join the bank stream (after KB_DEVICE_PTR calls) */
int kb_fx_bank_set_stream(kb_fx_bank* bank, void* cuda_stream);   /* run on a caller-owned cudaStream_t (NULL = bank's own) */
/* algorithmic (unique) HBM bytes one frame of one instance moves with the current controls */
double kb_fx_bank_bytes_per_frame(kb_fx_bank* bank);
long long kb_fx_bank_launches(const kb_fx_bank* bank);   /* kernels launched so far */
/* how many instances the last process() ran on the chunk-parallel schedule (the others ran frame-sequentially because
 * their control smoothers were still moving or their delays were shorter than a useful chunk); joins the stream */
int kb_fx_bank_parallel_instances(kb_fx_bank* bank);
/* how many instances the last process(KB_FX_TOLERANCE) ran on a re-associating schedule (0 when the flag was not given) */
int kb_fx_bank_tolerance_instances(kb_fx_bank* bank);
long long kb_fx_bank_state_bytes(const kb_fx_bank* bank); /* bytes of instance state mirrored between host and device */
/* Measurement: when enabled, every process() brackets its dominant kernel with CUDA events on the bank stream;
 * read() joins the stream and returns the accumulated kernel milliseconds and launch count since enable. */
int kb_fx_bank_profile(kb_fx_bank* bank, int enable);
int kb_fx_bank_profile_read(kb_fx_bank* bank, double* kernel_ms, long long* kernel_launches);

/* Factory presets of the bound programs (Plugin::presets, klang.h:1940-1981, 4195-4200; e.g. examples/PingPong.k:23-32).
 * kb_graph_preset returns the number of values of preset `index` of the program behind `graph` (is_synth selects KB_SY_* / KB_FX_*) and copies
 * its name and values.  load_preset is what a host does with one: every value through Control::set (klang.h:1725-1728), then onPreset. */
int kb_graph_num_presets(int is_synth, int graph);
int kb_graph_preset(int is_synth, int graph, int index, char* name, int name_max, float* values, int max_values);
int kb_fx_bank_load_preset(kb_fx_bank* bank, int instance, int index);

/* Debug taps: `x >> debug` in a program's process() (klang.h:3132-3287 Debug / Debug::Buffer / Debug::Session; examples/PingPong.k:61,
 * Gain/RM.k:22, Gain/Tremolo.k:27, Modulation/ModDelay.k:24).  In the reference the host opens a Debug::Session per block (clears the buffer),
 * every `>> debug` adds into the current frame's sample, and Session::getAudio() hands the block's capture to the IDE.  Here: while enabled,
 * each process() also leaves the capture of every instance on the device; debug_read copies the last block's capture, dst = [instances][n]
 * (host memory, or device memory with KB_DEVICE_PTR), and returns 1 — or 0 if the bank's program has no tap / nothing new (Buffer::get). */
int kb_fx_bank_debug_enable(kb_fx_bank* bank, int enable);
int kb_fx_bank_debug_read(kb_fx_bank* bank, float* dst, int n, unsigned flags);

/* ------------------------------------------------------------------------------------- synth bank */
/* `instances` Synth objects of `graph`, each with `voices` notes (notes.add<MyNote>(voices), <= 128, klang.h:4311). */
kb_synth_bank* kb_synth_bank_create(int graph, int instances, int voices, float fs, int max_block, int device);
void kb_synth_bank_destroy(kb_synth_bank* bank);
int kb_synth_bank_channels(const kb_synth_bank* bank);
int kb_synth_bank_instances(const kb_synth_bank* bank);
int kb_synth_bank_voices(const kb_synth_bank* bank);
int kb_synth_bank_num_controls(const kb_synth_bank* bank);
int kb_synth_bank_set_control(kb_synth_bank* bank, int instance, int idx, float value);
int kb_synth_bank_get_control(kb_synth_bank* bank, int instance, int idx, float* value);
/* Synth::noteOn(pitch, velocity): Notes::assign() voice allocation / stealing, then Note::start().
 * Returns the voice index (>= 0) or a negative error.                                    klang.h:4423-4427, 4336-4372 */
int kb_synth_bank_note_on(kb_synth_bank* bank, int instance, int pitch, float velocity);
/* Synth::noteOff(pitch, velocity): releases every sustaining note of that pitch           klang.h:4430-4434 */
int kb_synth_bank_note_off(kb_synth_bank* bank, int instance, int pitch, float velocity);
/* Synth::input(status, byte1, byte2), the raw MIDI entry of the v0.7.2 template: 0x90 with velocity > 0 = noteOn(byte1, byte2 / 127.f),
 * 0x80 or 0x90 with velocity 0 = noteOff, anything else = onMIDI() (a no-op for the bound graphs).  Returns 0 or a negative
 * error.                                                                 templates/juce/synth/Source/klang.h:3921-3929 */
int kb_synth_bank_midi(kb_synth_bank* bank, int instance, int status, int byte1, int byte2);
/* Synth::onControl(index, value) / Synth::onPreset(index): the synth's control() / preset() hook, then the hook of every note that is not Off
 * (klang.h:4399-4404, 4415-4420; Stereo::Synth 4789-4810).  No bound program overrides these hooks, so the audio is unaffected; the calls return
 * how many notes the reference would notify (stages as the device has evolved them).  load_preset = the preset's values through Control::set, then
 * onPreset.  (Synth::onMIDI, klang.h:4407-4412, calls itself and cannot return in the reference; raw MIDI enters through kb_synth_bank_midi.) */
int kb_synth_bank_on_control(kb_synth_bank* bank, int instance, int idx, float value);
int kb_synth_bank_on_preset(kb_synth_bank* bank, int instance, int index);
int kb_synth_bank_load_preset(kb_synth_bank* bank, int instance, int index);
/* NoteBase::start / release / stage of one voice                                          klang.h:4257-4284 */
int kb_synth_bank_voice_start(kb_synth_bank* bank, int instance, int voice, float pitch, float velocity);
int kb_synth_bank_voice_release(kb_synth_bank* bank, int instance, int voice, float velocity);
int kb_synth_bank_voice_stage(kb_synth_bank* bank, int instance, int voice);   /* 0 Onset 1 Sustain 2 Release 3 Off, <0 error */
/* A block's worth of events in one call, applied in order (what a host's MIDI loop does between process() calls,
 * templates/juce/synth/Source/PluginProcessor.cpp:169-174).  key = MIDI pitch for NOTE_ON/NOTE_OFF, voice index for
 * VOICE_START/VOICE_RELEASE, control index for CONTROL (value in `velocity`). */
#define KB_EV_NOTE_ON 0
#define KB_EV_NOTE_OFF 1
#define KB_EV_VOICE_START 2
#define KB_EV_VOICE_RELEASE 3
#define KB_EV_CONTROL 4
typedef struct kb_note_event { int type, instance, key; float pitch, velocity; } kb_note_event;
int kb_synth_bank_events(kb_synth_bank* bank, int count, const kb_note_event* events);
/* kb_synth_bank_events followed by kb_synth_bank_process in one call (a host's audio callback: MIDI loop, then the block) */
int kb_synth_bank_step(kb_synth_bank* bank, int count, const kb_note_event* events, float* out, int n, unsigned flags);
/* Synth::process(float*, int) / Stereo::Synth::process(float**, int) for every instance.
 * out = [instances][channels][n] (see flags for the other layouts); out is overwritten.   klang.h:4440-4466, 4830-4858 */
int kb_synth_bank_process(kb_synth_bank* bank, float* out, int n, unsigned flags);
int kb_synth_bank_sync(kb_synth_bank* bank);
int kb_synth_bank_set_stream(kb_synth_bank* bank, void* cuda_stream);
long long kb_synth_bank_launches(const kb_synth_bank* bank);
long long kb_synth_bank_state_bytes(const kb_synth_bank* bank);
/* bytes moved over PCIe by this bank so far: event state uploads (h2d), state fetches and host-buffer outputs (d2h) */
int kb_synth_bank_transfer_bytes(const kb_synth_bank* bank, long long* h2d, long long* d2h);
int kb_synth_bank_profile(kb_synth_bank* bank, int enable);
int kb_synth_bank_profile_read(kb_synth_bank* bank, double* kernel_ms, long long* kernel_launches);

/* --------------------------------------------------------------------- multi-GPU mix-down (one node) */
/* The polyphonic mix-down of a bank sharded over the GPUs of one box (the path's only exchange; it replaces
 * nothing in klang.h, which runs one Synth per host thread and sums on the host, klang.h:4842-4848).  One process per
 * GPU.  Rank 0 owns an arena of [2][world][max_floats] slots; every other rank maps it over NVLink (CUDA IPC) and passes
 * its slot as the `out` of kb_synth_bank_process(KB_DEVICE_PTR | KB_BANK_MIX): the bank-mix kernel itself stores the
 * rank's [channels][n] mix into rank 0's memory (no staging copy, no collective library on the data path).
 * publish() then raises the rank's flag; collect() on rank 0 waits for the flags of the step and sums the slots in RANK
 * ORDER (deterministic fp32).  Slots are double buffered by step parity; acquire() holds a rank that is two steps ahead.
 * Per step and rank: p = acquire(); process(out = p); publish(); rank 0 additionally: collect(dst). */
typedef struct kb_mixdown kb_mixdown;
#define KB_IPC_HANDLE_BYTES 64
kb_mixdown* kb_mixdown_create(int device, int world, int rank, int max_floats);
void kb_mixdown_destroy(kb_mixdown* m);
/* rank 0: the CUDA IPC handle of the arena (KB_IPC_HANDLE_BYTES bytes), to be sent to the other ranks by any means */
int kb_mixdown_export(kb_mixdown* m, void* handle);
/* ranks > 0: map rank 0's arena */
int kb_mixdown_import(kb_mixdown* m, const void* handle);
/* device pointer of this rank's slot for the next step (peer memory on ranks > 0); on `cuda_stream` the slot is first
 * waited free (rank 0 has consumed the step before the previous one).  NULL on error. */
float* kb_mixdown_acquire(kb_mixdown* m, void* cuda_stream);
/* after the kernels that filled the slot, on the same stream: make it visible to rank 0 and raise this rank's flag */
int kb_mixdown_publish(kb_mixdown* m, void* cuda_stream);
/* acquire + copy `count` floats of device memory into the slot + publish, for mixes that were finished locally */
int kb_mixdown_put(kb_mixdown* m, const float* src, int count, void* cuda_stream);
/* rank 0: dst[i] = slot[0][i] + slot[1][i] + ... (rank order) for i < count, once every rank has published this step */
int kb_mixdown_collect(kb_mixdown* m, float* dst, int count, void* cuda_stream);
/* The fused form — ONE kernel per block and rank, no acquire / publish / collect launches: the kernel waits for the slot of the step to be
 * free, stores `count` floats of `src` (device memory) into it, raises the rank's flag, and on rank 0 additionally sums the slots of the
 * PREVIOUS step in rank order into out_prev (device memory, or page-locked device-mapped host memory; rank 0 only; ignored elsewhere).  The exchange of block k thus runs beside the
 * kernels of block k + 1; kb_mixdown_collect() after the last block returns the last sum.  Do not mix with acquire / publish in one step. */
int kb_mixdown_step(kb_mixdown* m, const float* src, int count, float* out_prev, void* cuda_stream);
/* queue on `stream` a wait for the exchange kernel of the last fused step (a consumer of out_prev on a stream of its own) */
int kb_mixdown_stream_wait(kb_mixdown* mix, void* stream);
/* kb_synth_bank_process(KB_BANK_MIX) whose bank-mix kernel IS that fused step: the in-order sum of the bank's instances goes straight into
 * rank 0's arena over NVLink (Synth voices / instances shard across GPUs; this is the path's only exchange). */
int kb_synth_bank_process_mixdown(kb_synth_bank* bank, kb_mixdown* m, float* out_prev, int n, unsigned flags);
/* kb_synth_bank_events + kb_synth_bank_process_mixdown in one call (what a host's audio callback does per block, klang.h:4399-4466 over a sharded bank) */
int kb_synth_bank_step_mixdown(kb_synth_bank* bank, int count, const kb_note_event* events, kb_mixdown* m, float* out_prev, int n, unsigned flags);
/* block the host until the exchange kernel of the fused step `back` steps before the last one (0 .. 3) has finished: its out_prev is then valid */
int kb_mixdown_host_wait(kb_mixdown* mix, int back);

/* ---------------------------------------------------------------------------- primitive operators */
/* Single-object runs of the primitive operators ON THE DEVICE (one thread), for known-answer tests.
 * kinds as in tests/cases.py.  Generators::Fast / Basic / Wavetables (klang.h:4893-5381).  Kinds 12 / 13 = Basic::Noise / Fast::Noise
 * (klang.h:4947-4951, 5357-5366): the n ticks are the next n draws of the PROCESS's libc rand() stream, produced on the device,
 * and the call leaves libc's rand() advanced by n draws exactly as the reference's loop would. */
int kb_prim_osc(int kind, int nargs, float f, float phase, float duty, float fs, int n, float* out);
/* File::WAV (klang.h:5951-6085: load + operator>> into a variable::buffer) over a file image in host memory — host code in the reference as
 * well; returns the number of decoded samples (data->size / BlockAlign), copies at most max_samples, info = { NumChannels, SampleRate,
 * BitsPerSample }.  klang::Sample (klang.h:3679-3720: set(f) [nargs 1] or set(f, phase) [nargs 2], then n ticks) played on the device from a
 * table in host memory, e.g. the decoded file. */
int kb_wav_decode(const void* image, long long nbytes, float* out, int max_samples, int* info);
int kb_prim_sample(const float* table, int size, int nargs, float f, float phase, int n, float* out);
/* One Delay<1000> (klang.h:3381-3512), sample by sample: write in[s]; out_i = tap(int di[s]); out_f = tap(float df[s]);
 * out_l = lagrange(df[s]) (klang.h:3429-3458); set(set_at[s]) when set_at[s] >= 0; out_p = process() once set() has placed a read
 * head, else 0.  Delays must lie inside the line (0 <= di < 1000, 0 <= df < 999). */
int kb_prim_delay(int n, const float* in, const int* di, const float* df, const float* set_at,
                  float* out_i, float* out_f, float* out_p, float* out_l);
/* Filters::Biquad::{LPF,HPF,BPF,BRF,APF}, OnePole::{LPF,HPF}, Butterworth::LPF<1>,<2>, DCF, IIR<1>, IIR<2>, Modifiers::Modal,
 * Envelope::Follower peak / rms, Follower::Window<64> mean / rms (klang.h:5387-5948; kinds as in tests/cases.py FLT_*):
 * set(f[s],Q[s]) before sample s < nset (one-pole, Modal and Follower kinds: nset <= 1, coefficients from the host libm;
 * DCF: f = r; IIR<1>: f = coefficient; IIR<2>: f = a1, Q = a2; Modal: f, Q = decay seconds; Follower / Window: f = attack,
 * Q = release seconds) */
int kb_prim_filter(int kind, int nset, const float* f, const float* Q, float fs, int n, const float* in, float* out, float* coeffs);
/* Envelope / ADSR (klang.h:3723-4137) */
int kb_prim_envelope(int npts, const float* xy, int loop_start, int loop_end, float fs, int n, int release_at,
                     float release_time, float release_level, float* out, int* stage_out);
int kb_prim_adsr(float A, float D, float S, float R, float fs, int n, int release_at, float* out, int* stage_out);
/* Wavetables::Sine / Saw (kind 10 / 11, klang.h:5372-5379): the 2048-entry table as the device oscillators read it (filled by the
 * constructor code of Wavetable::operator=(Oscillator), klang.h:3645-3650, on the host; resident in HBM) */
int kb_prim_wavetable(int kind, float fs, float* out2048);
/* One Stereo::Delay<1000> (klang.h:4647-4700), sample by sample: write {inl[s], inr[s]}; {outl, outr}[s] = tap(float df[s]) */
int kb_prim_stereo_delay(int n, const float* inl, const float* inr, const float* df, float* outl, float* outr);
/* Control::set(values[s]) then Control::smooth() per sample on a Dial(lo, hi, initial)       klang.h:1715-1728, 1796-1799 */
int kb_prim_control_smooth(float lo, float hi, float initial, int n, const float* values, float* out);
/* Envelope::at(t[s]) on the breakpoints xy = {x0, y0, x1, y1, ...}                              klang.h:3929-3942 */
int kb_prim_envelope_at(int npts, const float* xy, int n, const float* t, float* out);
/* libm agreement probe: out[i] = device sinf / cosf / tanhf / expf of x[i] (fn 0 / 1 / 2 / 3) */
int kb_prim_math(int fn, int n, const float* x, float* out);

#ifdef __cplusplus
}
#endif
#endif
