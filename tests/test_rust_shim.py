"""The Rust shim (rust/oddio-b200) cannot be compiled here (no rustc / cargo in the image), so its `extern "C"`
block is checked against include/oddio_b200.h textually: every entry point of the header is declared exactly once,
with the same number of arguments, and every argument / return type maps to the C type the header states."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "oddio_b200.h")
LIB_RS = os.path.join(ROOT, "rust", "oddio-b200", "src", "lib.rs")

C_TO_RUST = {
    "int": "c_int", "uint32_t": "u32", "uint64_t": "u64", "float": "f32", "double": "f64", "odb_frames": "odb_frames",
    "odb_source": "odb_source", "const char*": "*const c_char", "void*": "*mut c_void", "const void*": "*const c_void",
    "void**": "*mut *mut c_void", "float*": "*mut f32", "const float*": "*const f32", "double*": "*mut f64", "int*": "*mut c_int",
    "uint32_t*": "*mut u32", "uint64_t*": "*mut u64", "int16_t*": "*mut i16", "const int16_t*": "*const i16",
    "const uint8_t*": "*const u8", "odb_frames*": "*mut odb_frames", "odb_source*": "*mut odb_source",
    "const odb_source*": "*const odb_source", "const odb_chain*": "*const odb_chain",
    "odb_ctx*": "*mut odb_ctx", "odb_ctx**": "*mut *mut odb_ctx", "odb_scene*": "*mut odb_scene", "odb_scene**": "*mut *mut odb_scene",
    "odb_mixer*": "*mut odb_mixer", "odb_mixer**": "*mut *mut odb_mixer", "odb_exchange*": "*mut odb_exchange",
    "odb_exchange**": "*mut *mut odb_exchange",
}


def c_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(odb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = re.sub(r"\[\d*\]", "*", a.strip())       # float x[3] -> float x*
                a = re.sub(r"\s+", " ", a)
                mm = re.match(r"(.*?)([A-Za-z_]\w*)?(\*?)$", a)  # strip the parameter name
                ty = (mm.group(1) + mm.group(3)).strip() if mm else a
                ty = ty.replace(" *", "*").replace("* ", "*").strip()
                params.append(ty)
        out[name] = (ret.replace(" *", "*"), params)
    return out


def rust_decls():
    src = open(LIB_RS).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n    }\n")]
    out = {}
    for m in re.finditer(r"pub fn (odb_[a-z0-9_]+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        name, args, ret = m.group(1), m.group(2).strip(), (m.group(3) or "").strip()
        params = [a.split(":", 1)[1].strip() for a in args.split(",") if a.strip()] if args else []
        assert name not in out, f"{name} declared twice"
        out[name] = (ret, params)
    return out


def test_extern_block_matches_the_header():
    c, r = c_decls(), rust_decls()
    assert len(c) >= 56
    assert set(c) == set(r), f"only in the header: {sorted(set(c) - set(r))}; only in lib.rs: {sorted(set(r) - set(c))}"
    for name, (cret, cparams) in c.items():
        rret, rparams = r[name]
        assert C_TO_RUST[cret] == rret, f"{name}: returns {cret} in C, {rret} in Rust"
        assert len(cparams) == len(rparams), f"{name}: {len(cparams)} parameters in C, {len(rparams)} in Rust"
        for i, (ct, rt) in enumerate(zip(cparams, rparams)):
            assert ct in C_TO_RUST, f"{name} parameter {i}: unmapped C type {ct!r}"
            assert C_TO_RUST[ct] == rt, f"{name} parameter {i}: {ct} in C, {rt} in Rust"


def test_chain_struct_layout_matches():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = h[h.index("typedef struct odb_chain {"):h.index("} odb_chain;")]
    c_fields = re.findall(r"(\w+)\s+(\w+);", body)
    r = open(LIB_RS).read()
    rb = r[r.index("pub struct odb_chain {"):]
    rb = rb[:rb.index("}")]
    r_fields = re.findall(r"pub (\w+): (\w+),", rb)
    assert [(C_TO_RUST[t], n) for t, n in c_fields] == [(t, n) for n, t in r_fields]
    assert "#[repr(C)]" in r[r.index("pub struct odb_chain {") - 80:r.index("pub struct odb_chain {")]


def test_shim_keeps_the_reference_surface():
    """The names a user of the reference calls exist with the reference's shapes."""
    r = open(LIB_RS).read()
    for needle in ("pub trait Signal", "pub trait Seek: Signal", "pub trait Frame", "pub fn run<S: Signal + ?Sized>(signal: &mut S, sample_rate: u32",
                   "impl Signal for SpatialScene", "impl<T: DeviceFrame> Signal for Mixer<T>", "pub fn new(ctx: &Arc<Context>) -> (SpatialSceneControl, Self)",
                   "pub fn play<S: DeviceSeek<Frame = Sample>>(&mut self, mut signal: S, options: SpatialOptions) -> Spatial",
                   "pub fn play_buffered<S: DeviceSignal<Frame = Sample>>", "pub fn set_listener_rotation(&mut self, rotation: mint::Quaternion<f32>)",
                   "pub fn set_motion(&mut self, position: mint::Point3<f32>, velocity: mint::Vector3<f32>, discontinuity: bool)",
                   "pub fn is_finished(&self) -> bool", "pub fn stop(&mut self)", "pub fn is_stopped(&self) -> bool",
                   "pub struct Tanh<A>", "pub struct Reinhard<A>", "pub struct Cycle<T: DeviceFrame>", "pub fn set_speed(&mut self, factor: f32)",
                   "pub fn set_amplitude_ratio(&mut self, factor: f32)", "pub fn playback_position(&self) -> f64"):
        assert needle in r, f"missing from the shim: {needle}"
    # Speed and Gain are not Seek in the reference (speed.rs:26-40, gain.rs:95-127): no DeviceSeek impl for them
    assert "DeviceSeek for Speed" not in r and "DeviceSeek for Gain<" not in r
