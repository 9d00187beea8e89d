"""kcc — translate a klang `.k` Effect program and compile its own process() body for the B200 (Tier B, first step).

  python -m klang_b200.kcc examples/Gain/Gain.k -o /tmp/libgain_k.so

The translation is deliberately small and textual (klang programs are plain C++ over the klang.h DSL):
  * `#include <klang.h>` is replaced by the device-capable subset klang_b200/csrc/kb_kdev.cuh;
  * every function DEFINITION of the file — free functions and member functions other than constructors — is marked __host__ __device__;
  * `virtual` / `override` are dropped (the kernels call process() on the concrete type; an object with a host vptr could not travel);
  * the plugin (the struct derived from Effect / Stereo::Effect) is exported behind include/klang_b200_user.h.
Everything else — the controls table, the body of process(), helper functions — is compiled exactly as written."""
import argparse
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_KEYWORDS = {"if", "else", "for", "while", "switch", "return", "do", "catch", "sizeof", "new", "delete", "case"}
NVCC_FLAGS = ["-std=c++17", "-O3", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true", "-cudart", "static", "--expt-relaxed-constexpr"]


class KccError(RuntimeError):
    pass


def translate(text, source_name="program.k"):
    """Returns (cuda source, plugin name, channels)."""
    m = re.search(r"struct\s+(\w+)\s*:\s*(?:public\s+)?((?:Stereo::|stereo::)?Effect)\b", text)
    sm = re.search(r"struct\s+(\w+)\s*:\s*(?:public\s+)?Synth\b", text)
    note = None
    if sm:
        # a mono Synth: the note type is the one its constructor adds (notes.add<NOTE>(count), klang.h:4325-4331)
        nm = re.search(r"notes\s*\.\s*add\s*<\s*(\w+)\s*>\s*\(", text)
        if not nm:
            raise KccError(f"{source_name}: the synth adds no notes (notes.add<NOTE>(n))")
        plugin, channels, note = sm.group(1), 1, nm.group(1)
    elif m:
        plugin, channels = m.group(1), 2 if "::" in m.group(2) else 1
    else:
        raise KccError(f"{source_name}: no struct derived from Effect / Stereo::Effect / Synth (Stereo::Synth programs are not translated yet)")
    # Function<Args...> members hold a std::function in the reference (a host address).  The function each one is constructed with is read
    # from the constructor's initialiser list and named in the member's type: `Function<float, float> f;` + `Shaping() : f(softclip)`
    # become `klang::FunctionT<kb_fn_f, 2> f;` + `Shaping() : f()`, with `struct kb_fn_f { call(...) -> softclip(...) }` in front of the plugin.
    functors = []
    for fmem in re.finditer(r"\bFunction\s*<([^<>;]*)>\s+(\w+)\s*;", text):
        nargs, member = len([a for a in fmem.group(1).split(",") if a.strip()]), fmem.group(2)
        init = re.search(r"\)\s*:\s*(?:[^{;]*,\s*)?" + member + r"\s*\(\s*([A-Za-z_]\w*)\s*\)", text)
        if not init:
            raise KccError(f"{source_name}: Function member `{member}` is not constructed from a named function in a constructor initialiser")
        functors.append((member, nargs, init.group(1), fmem.group(0), init.group(0)))
    for member, nargs, fn, decl, init in functors:
        text = text.replace(decl, f"klang::FunctionT<kb_fn_{member}, {nargs}> {member};")
        text = text.replace(init, re.sub(r"\b" + member + r"\s*\(\s*" + fn + r"\s*\)$", f"{member}()", init))
    out = []
    depth = 0
    for lineno, line in enumerate(text.splitlines(), 1):
        code, sep, comment = line.partition("//")
        # a namespace-scope constant (`const float FREQ[6] = { ... };`) is a host object: device code gets a __device__ twin and the name picks
        # the one of the side it is compiled for
        gm = re.match(r"^\s*(?:static\s+)?const\s+([A-Za-z_][\w:]*)\s+([A-Za-z_]\w*)\s*((?:\[[^\]]*\])*)\s*=\s*(.+);\s*$", code) if depth == 0 else None
        depth += code.count("{") - code.count("}")
        if gm:
            ty, name, dims, init = gm.groups()
            out.append(f"const {ty} {name}_kbh{dims} = {init}; __device__ const {ty} {name}_kbd{dims} = {init};{sep}{comment}")
            out.append(f"#ifdef __CUDA_ARCH__\n#define {name} {name}_kbd\n#else\n#define {name} {name}_kbh\n#endif\n#line {lineno + 1}")
            continue
        # klang::fs is a host global and `debug` a host object: device code reads the bank's rate through kb_fs() and taps into a sink value
        code = re.sub(r"\bklang::fs\b", "kb_fs()", code)
        code = re.sub(r"(?<![\w.>:])fs\b(?!\s*[\(:])", "kb_fs()", code)
        code = re.sub(r">>\s*debug\b", ">> klang::Debug()", code)
        code = re.sub(r"(?<![\w.>:])graph\s*\.", "kb_graph().", code)             # `graph.clear()` / `graph.add(y)`: the UI plot, a host object there
        code = re.sub(r">>\s*graph\s*;", ">> kb_graph();", code)
        if functors and re.match(r"\s*struct\s+" + re.escape(plugin) + r"\b", code):
            args = ", ".join(f"float x{i}" for i in range(3))
            for member, nargs, fn, _, _ in functors:
                params = ", ".join(f"float x{i}" for i in range(nargs))
                call = ", ".join(f"x{i}" for i in range(nargs))
                out.append(f"struct kb_fn_{member} {{ KB_KD static float call({params}) {{ return {fn}({call}); }} }};")
                out.append(f"#line {lineno}")
        code = re.sub(r"(?<![\w.>:])(pi|ln2|root2)\b(?!\s*[\(:])", r"kb_\1()", code)
        code = re.sub(r"(?<![\w.>:])(min|max)\s*\(", r"kb_\1(", code)          # klang's own min / max (klang.h:221-224), not ::min / ::max
        code = re.sub(r"(?<![\w.>:])(tanh|exp|abs)\s*\(", r"kb_\1(", code)         # the float overloads that restate the host's libm on the device
        line = code + sep + comment
        if re.match(r"\s*#\s*include\s*<klang\.h>", line):
            out.append("// (klang.h -> klang_b200/csrc/kb_kdev.cuh)")
            continue
        # a program's own generator: the dataflow protocol dispatches statically, so the base learns the concrete type
        line = re.sub(r"\bstruct\s+(\w+)\s*:\s*(?:public\s+)?(Oscillator|Generator|Modifier)\b", r"struct \1 : \2T<\1>", line)
        line = re.sub(r"\bvirtual\s+", "", line)
        line = re.sub(r"\)\s*override\b", ")", line)
        # a function definition: [qualifiers] type name(args) [const] {     — not a control statement, not a constructor (no return type)
        fm = re.match(r"^(\s*)((?:static\s+|inline\s+|constexpr\s+)*)([A-Za-z_][\w:<>,\*&\s]*?[\w>\*&])\s+([A-Za-z_]\w*)\s*\(([^;{}]*)\)\s*(const\s*)?\{?\s*$", line)
        # a note's on() / off() run on the host mirror only (kb_kdev.cuh): they stay host functions, so they may read host tables (FM.k)
        if fm and fm.group(3).strip() == "event" and fm.group(4) in ("on", "off"):
            fm = None
        if fm and fm.group(4) not in _KEYWORDS and fm.group(3).split()[-1] not in _KEYWORDS | {"struct", "class", "namespace", "using", "typedef"} \
                and "=" not in fm.group(3):
            line = f"{fm.group(1)}KB_KD {re.sub(r'^((?:static\s+|constexpr\s+)*)inline\s+', r'\1', line.lstrip())}"   # (KB_KD carries the inline)
        out.append(line)
    body = "\n".join(out)
    src = (f'// GENERATED by klang_b200/kcc.py from {source_name}: the program text below is the user\'s, with its functions marked __host__ __device__\n'
           f'#include "kb_kdev.cuh"\n'
           f'namespace kb_user {{\nusing namespace klang;\n#line 1 "{source_name}"\n{body}\n}}\n'
           + (f'KB_USER_EXPORT_SYNTH(kb_user::{plugin}, kb_user::{plugin}::{note}, "{plugin}")\n' if note else f'KB_USER_EXPORT(kb_user::{plugin}, "{plugin}")\n'))
    return src, plugin, channels


def compile_k(path, out, keep_source=None, verbose=False):
    """Translate `path` and build the shared library `out` with nvcc (cross-compiles without a GPU).  Returns out."""
    with open(path, encoding="utf-8", errors="replace") as f:
        src, plugin, channels = translate(f.read(), os.path.basename(path))
    cu = keep_source or (os.path.splitext(out)[0] + ".cu")
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    with open(cu, "w") as f:
        f.write(src)
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(_HERE, "csrc"), cu, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise KccError(f"nvcc failed for {path} (translated source kept in {cu})")
    return out


class UserFx:
    """A bank of `instances` objects of a translated `.k` Effect (include/klang_b200_user.h), with FxBank's call surface."""

    def __init__(self, so_path, instances=1, fs=44100.0, max_block=16384, device=0):
        L = C.CDLL(so_path)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        for name, res, args in (("kb_user_name", C.c_char_p, []), ("kb_user_channels", i, []), ("kb_user_num_controls", i, []), ("kb_user_stateless", i, []),
                                ("kb_user_last_error", C.c_char_p, []), ("kb_user_fx_create", vp, [i, f, i, i]), ("kb_user_fx_destroy", None, [vp]),
                                ("kb_user_fx_set_control", i, [vp, i, i, f]), ("kb_user_fx_get_control", i, [vp, i, i, C.POINTER(f)]),
                                ("kb_user_fx_process", i, [vp, vp, i, C.c_uint])):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.L = L
        self.name, self.channels, self.num_controls = L.kb_user_name().decode(), L.kb_user_channels(), L.kb_user_num_controls()
        self.stateless = bool(L.kb_user_stateless())
        self.instances = instances
        self.h = L.kb_user_fx_create(instances, float(fs), max_block, device)
        if not self.h:
            raise KccError("kb_user_fx_create: " + L.kb_user_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.kb_user_fx_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc < 0:
            raise KccError(f"{what}: error {rc}: {self.L.kb_user_last_error().decode()}")

    def set_control(self, idx, value, instance=None):
        for k in (range(self.instances) if instance is None else [instance]):
            self._check(self.L.kb_user_fx_set_control(self.h, k, idx, float(value)), "kb_user_fx_set_control")

    def get_control(self, idx, instance=0):
        v = C.c_float()
        self._check(self.L.kb_user_fx_get_control(self.h, instance, idx, C.byref(v)), "kb_user_fx_get_control")
        return float(v.value)

    def process_inplace(self, io, n=None):
        """io: float32 [instances, channels, n], numpy (host) or torch CUDA tensor."""
        n = io.shape[-1] if n is None else n
        if isinstance(io, np.ndarray):
            if io.dtype != np.float32 or not io.flags["C_CONTIGUOUS"] or io.size < self.instances * self.channels * n:
                raise KccError("process_inplace: need a C-contiguous float32 array [instances, channels, n]")
            p, dev = io.ctypes.data, 0
        else:
            if str(io.dtype) != "torch.float32" or not io.is_contiguous() or io.numel() < self.instances * self.channels * n or not io.is_cuda:
                raise KccError("process_inplace: need a contiguous float32 CUDA tensor [instances, channels, n]")
            p, dev = io.data_ptr(), 1
        self._check(self.L.kb_user_fx_process(self.h, p, n, dev), "kb_user_fx_process")
        return io


class UserSynth:
    """A bank of `instances` objects of a translated `.k` Synth (klang_b200/csrc/kb_kdev.cuh: kb_user_synth_*), one voice per note the program
    adds; the call surface of SynthBank for one instance at a time."""

    def __init__(self, so_path, instances=1, fs=44100.0, max_block=16384, device=0):
        L = C.CDLL(so_path)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        for name, res, args in (("kb_user_name", C.c_char_p, []), ("kb_user_kind", i, []), ("kb_user_num_controls", i, []), ("kb_user_synth_voices", i, []),
                                ("kb_user_last_error", C.c_char_p, []), ("kb_user_synth_create", vp, [i, f, i, i]), ("kb_user_synth_destroy", None, [vp]),
                                ("kb_user_synth_set_control", i, [vp, i, i, f]), ("kb_user_synth_get_control", i, [vp, i, i, C.POINTER(f)]),
                                ("kb_user_synth_voice_start", i, [vp, i, i, f, f]), ("kb_user_synth_voice_release", i, [vp, i, i, f]),
                                ("kb_user_synth_voice_stage", i, [vp, i, i]), ("kb_user_synth_note_on", i, [vp, i, i, f]), ("kb_user_synth_note_off", i, [vp, i, i, f]),
                                ("kb_user_synth_process", i, [vp, vp, i, C.c_uint])):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.L = L
        if L.kb_user_kind() != 1:
            raise KccError(f"{so_path}: not a translated Synth")
        self.name, self.num_controls, self.voices = L.kb_user_name().decode(), L.kb_user_num_controls(), L.kb_user_synth_voices()
        self.instances, self.channels = instances, 1
        self.h = L.kb_user_synth_create(instances, float(fs), max_block, device)
        if not self.h:
            raise KccError("kb_user_synth_create: " + L.kb_user_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.kb_user_synth_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc < 0:
            raise KccError(f"{what}: error {rc}: {self.L.kb_user_last_error().decode()}")
        return rc

    def set_control(self, idx, value, instance=None):
        for k in (range(self.instances) if instance is None else [instance]):
            self._check(self.L.kb_user_synth_set_control(self.h, k, idx, float(value)), "kb_user_synth_set_control")

    def get_control(self, idx, instance=0):
        v = C.c_float()
        self._check(self.L.kb_user_synth_get_control(self.h, instance, idx, C.byref(v)), "kb_user_synth_get_control")
        return float(v.value)

    def voice_start(self, voice, pitch, velocity, instance=0):
        self._check(self.L.kb_user_synth_voice_start(self.h, instance, voice, float(pitch), float(velocity)), "kb_user_synth_voice_start")

    def voice_release(self, voice, velocity=0.0, instance=0):
        self._check(self.L.kb_user_synth_voice_release(self.h, instance, voice, float(velocity)), "kb_user_synth_voice_release")

    def voice_stage(self, voice, instance=0):
        return self._check(self.L.kb_user_synth_voice_stage(self.h, instance, voice), "kb_user_synth_voice_stage")

    def note_on(self, pitch, velocity, instance=0):
        return self._check(self.L.kb_user_synth_note_on(self.h, instance, int(pitch), float(velocity)), "kb_user_synth_note_on")

    def note_off(self, pitch, velocity=0.0, instance=0):
        self._check(self.L.kb_user_synth_note_off(self.h, instance, int(pitch), float(velocity)), "kb_user_synth_note_off")

    def process_block(self, n, per_voice=False):
        """float32 [instances, n] (Synth::process), or [instances * voices, n] per-voice streams."""
        out = np.zeros((self.instances * (self.voices if per_voice else 1), n), np.float32)
        self._check(self.L.kb_user_synth_process(self.h, out.ctypes.data, n, 2 if per_voice else 0), "kb_user_synth_process")
        return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("k_file")
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--keep-source", default=None, help="where to leave the translated .cu")
    args = ap.parse_args()
    print(compile_k(args.k_file, args.out, args.keep_source, verbose=True))


if __name__ == "__main__":
    main()
