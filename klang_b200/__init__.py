"""klang_b200 — B200-native implementation of klang's per-sample signal-flow path.

Host-side mirror of the reference's plugin interface (`Effect` / `Synth` block drivers, klang.h:4203-4467,
4703-4859) over the C-ABI library klang_b200/lib/libklang_b200.so (include/klang_b200.h).  Everything that
processes samples runs in the library's sm_100a kernels; this package only marshals buffers.  Importing it
never touches oracle/.
"""
from .api import (  # noqa: F401
    Engine, FxBank, SynthBank, KlangB200Error, lib, lib_path, device_count, presets, wav_decode,
    FX_GAIN, FX_PINGPONG, FX_REVERB, FX_DELAY_PINGPONG, FX_DELAY_REVERB, FX_PAN, FX_RM, FX_TREMOLO, FX_CLIPPING, FX_ECHO, FX_FEEDBACK, FX_FUNCTIONS, FX_MUTE, FX_IIR, FX_WAHWAH, FX_FLANGER, FX_MODDELAY, FX_MOD_CHORUS,
    SY_SUBTRACTIVE, SY_SUPERSAW, SY_TB303, SY_SYNTHX, SY_FILTER_K, SY_FM, SY_BREAKPOINT, SY_RAMP, SY_RELEASE, SY_ADDITIVE_SAW, SY_ADDITIVE_SQUARE, SY_AM, SY_MOD_FM, SY_MOD_FM2, SY_ADDITIVE_NYQUIST,
    DEVICE_PTR, PER_VOICE, MIX_SUM, BANK_MIX, LANE_PER_VOICE, FX_SEQUENTIAL, ASYNC_HOST, FX_TOLERANCE,
    EV_NOTE_ON, EV_NOTE_OFF, EV_VOICE_START, EV_VOICE_RELEASE, EV_CONTROL, EVENT_DTYPE,
)

__version__ = "0.1.0"
