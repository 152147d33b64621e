//! oddio's hot path on a B200: the reference's `Signal` / `Seek` / `Frame` traits and the
//! `SpatialScene::new` / `play` / `set_motion` / `run` and `Mixer::new` / `play` / `stop` API, with the audio-thread
//! half (`Signal::sample` of `SpatialScene`, `Mixer`, `Tanh`, `Reinhard`) forwarded to `liboddio_b200.so` through the
//! C ABI of `include/oddio_b200.h`.
//!
//! The reference mixes `Box<dyn Signal>` (spatial.rs:14-15, mixer.rs:122); a device path can only run a closed set of
//! signal chains, so `play` is bounded by the sealed [`DeviceSignal`] trait: `[Gain]([FixedGain]([Speed](FramesSignal |
//! Cycle)))`. `SpatialSceneControl::play` additionally requires [`DeviceSeek`], which `Speed` and `Gain` do not
//! implement - exactly the reference's `Seek` bound (speed.rs:26-40, gain.rs:95-127).
//!
//! This crate cannot be compiled in the image the library is built in (no rustc); `tests/test_rust_shim.py` checks
//! its `extern "C"` block against the header, name by name and argument by argument.
#![allow(clippy::missing_safety_doc)]

use std::ffi::{c_char, c_int, c_void, CStr};
use std::marker::PhantomData;
use std::ptr;
use std::sync::Arc;

pub type Sample = f32;

// ---------------------------------------------------------------------------------------------------------------
// Raw bindings: one declaration per entry point of include/oddio_b200.h, in the header's order.
pub mod sys {
    use super::*;

    #[repr(C)]
    pub struct odb_ctx {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct odb_scene {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct odb_mixer {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct odb_exchange {
        _p: [u8; 0],
    }
    pub type odb_frames = u64;
    pub type odb_source = u64;

    pub const ODB_OK: c_int = 0;
    pub const ODB_E_INVALID: c_int = -1;
    pub const ODB_E_CUDA: c_int = -2;
    pub const ODB_E_UNSUPPORTED: c_int = -3;
    pub const ODB_E_NOMEM: c_int = -4;
    pub const ODB_CHAIN_SPEED: u32 = 0x1;
    pub const ODB_CHAIN_FIXED_GAIN: u32 = 0x2;
    pub const ODB_CHAIN_GAIN: u32 = 0x4;
    pub const ODB_CHAIN_CYCLE: u32 = 0x8;
    pub const ODB_EPILOGUE_NONE: c_int = 0;
    pub const ODB_EPILOGUE_TANH: c_int = 1;
    pub const ODB_EPILOGUE_REINHARD: c_int = 2;

    #[repr(C)]
    #[derive(Clone, Copy, Debug, Default)]
    pub struct odb_chain {
        pub frames: odb_frames,
        pub start_seconds: f64,
        pub flags: u32,
        pub speed: f32,
        pub fixed_gain_db: f32,
        pub gain_ratio: f32,
    }

    #[link(name = "oddio_b200")]
    extern "C" {
        pub fn odb_last_error() -> *const c_char;
        pub fn odb_abi_version() -> u32;
        pub fn odb_ctx_create(cuda_device: c_int, out: *mut *mut odb_ctx) -> c_int;
        pub fn odb_ctx_create_on_stream(cuda_device: c_int, cuda_stream: *mut c_void, out: *mut *mut odb_ctx) -> c_int;
        pub fn odb_ctx_destroy(ctx: *mut odb_ctx) -> c_int;
        pub fn odb_ctx_synchronize(ctx: *mut odb_ctx) -> c_int;
        pub fn odb_ctx_stream(ctx: *mut odb_ctx, out_stream: *mut *mut c_void) -> c_int;
        pub fn odb_pin_buffer(ctx: *mut odb_ctx, host_ptr: *mut c_void, bytes: u64) -> c_int;
        pub fn odb_unpin_buffer(ctx: *mut odb_ctx, host_ptr: *mut c_void) -> c_int;
        pub fn odb_frames_from_slice(ctx: *mut odb_ctx, rate: u32, channels: c_int, samples: *const f32, n_frames: u64, out: *mut odb_frames) -> c_int;
        pub fn odb_frames_from_device(ctx: *mut odb_ctx, rate: u32, channels: c_int, dev_samples: *const c_void, n_frames: u64, out: *mut odb_frames) -> c_int;
        pub fn odb_frames_from_i16(ctx: *mut odb_ctx, rate: u32, channels: c_int, samples: *const i16, n_frames: u64, bits_per_sample: c_int, out: *mut odb_frames) -> c_int;
        pub fn odb_frames_release(ctx: *mut odb_ctx, frames: odb_frames) -> c_int;
        pub fn odb_scene_create(ctx: *mut odb_ctx, out: *mut *mut odb_scene) -> c_int;
        pub fn odb_scene_destroy(scene: *mut odb_scene) -> c_int;
        pub fn odb_scene_set_epilogue(scene: *mut odb_scene, epilogue: c_int) -> c_int;
        pub fn odb_scene_play(scene: *mut odb_scene, chain: *const odb_chain, position: *const f32, velocity: *const f32, radius: f32, out: *mut odb_source) -> c_int;
        pub fn odb_scene_play_buffered(scene: *mut odb_scene, chain: *const odb_chain, position: *const f32, velocity: *const f32, radius: f32, max_distance: f32, rate: u32, buffer_duration: f32, out: *mut odb_source) -> c_int;
        pub fn odb_scene_set_listener_rotation(scene: *mut odb_scene, q_xyzs: *const f32) -> c_int;
        pub fn odb_spatial_set_motion(scene: *mut odb_scene, src: odb_source, position: *const f32, velocity: *const f32, discontinuity: c_int) -> c_int;
        pub fn odb_spatial_set_motion_many(scene: *mut odb_scene, n: u32, srcs: *const odb_source, positions: *const f32, velocities: *const f32, discontinuity: *const u8) -> c_int;
        pub fn odb_spatial_is_finished(scene: *mut odb_scene, src: odb_source, out: *mut c_int) -> c_int;
        pub fn odb_scene_sample(scene: *mut odb_scene, interval: f32, out: *mut f32, n_frames: u32) -> c_int;
        pub fn odb_scene_sample_i16(scene: *mut odb_scene, interval: f32, out: *mut i16, n_frames: u32) -> c_int;
        pub fn odb_scene_run(scene: *mut odb_scene, sample_rate: u32, out: *mut f32, n_frames: u32) -> c_int;
        pub fn odb_scene_sample_device(scene: *mut odb_scene, interval: f32, dev_out: *mut c_void, n_frames: u32) -> c_int;
        pub fn odb_scene_len(scene: *mut odb_scene, buffered: c_int, out: *mut u64) -> c_int;
        pub fn odb_mixer_create(ctx: *mut odb_ctx, channels: c_int, out: *mut *mut odb_mixer) -> c_int;
        pub fn odb_mixer_destroy(mixer: *mut odb_mixer) -> c_int;
        pub fn odb_mixer_set_epilogue(mixer: *mut odb_mixer, epilogue: c_int) -> c_int;
        pub fn odb_mixer_play(mixer: *mut odb_mixer, chain: *const odb_chain, out: *mut odb_source) -> c_int;
        pub fn odb_mixed_stop(mixer: *mut odb_mixer, src: odb_source) -> c_int;
        pub fn odb_mixed_is_stopped(mixer: *mut odb_mixer, src: odb_source, out: *mut c_int) -> c_int;
        pub fn odb_mixer_sample(mixer: *mut odb_mixer, interval: f32, out: *mut f32, n_frames: u32) -> c_int;
        pub fn odb_mixer_run(mixer: *mut odb_mixer, sample_rate: u32, out: *mut f32, n_frames: u32) -> c_int;
        pub fn odb_mixer_sample_i16(mixer: *mut odb_mixer, interval: f32, out: *mut i16, n_frames: u32) -> c_int;
        pub fn odb_mixer_sample_device(mixer: *mut odb_mixer, interval: f32, dev_out: *mut c_void, n_frames: u32) -> c_int;
        pub fn odb_mixer_len(mixer: *mut odb_mixer, out: *mut u64) -> c_int;
        pub fn odb_source_set_speed(owner: *mut c_void, src: odb_source, factor: f32) -> c_int;
        pub fn odb_source_set_amplitude_ratio(owner: *mut c_void, src: odb_source, factor: f32) -> c_int;
        pub fn odb_source_set_gain_db(owner: *mut c_void, src: odb_source, db: f32) -> c_int;
        pub fn odb_source_playback_position(owner: *mut c_void, src: odb_source, out_seconds: *mut f64) -> c_int;
        pub fn odb_source_frames_is_finished(owner: *mut c_void, src: odb_source, out: *mut c_int) -> c_int;
        pub fn odb_source_cursor(owner: *mut c_void, src: odb_source, out_t: *mut f64, out_ring_write: *mut f32) -> c_int;
        pub fn odb_last_launch_count(owner: *mut c_void, out: *mut u32) -> c_int;
        pub fn odb_last_job_counters(owner: *mut c_void, out: *mut u32) -> c_int;
        pub fn odb_set_profiling(owner: *mut c_void, enabled: c_int) -> c_int;
        pub fn odb_last_mix_kernel_ms(owner: *mut c_void, out_ms: *mut f32) -> c_int;
        pub fn odb_set_kernel_variant(owner: *mut c_void, variant: c_int) -> c_int;
        pub fn odb_exchange_create(ctx: *mut odb_ctx, rank: c_int, world: c_int, max_floats: u32, depth: c_int, out: *mut *mut odb_exchange) -> c_int;
        pub fn odb_exchange_destroy(ex: *mut odb_exchange) -> c_int;
        pub fn odb_exchange_handle_size() -> c_int;
        pub fn odb_exchange_export(ex: *mut odb_exchange, handle_out: *mut c_void) -> c_int;
        pub fn odb_exchange_connect(ex: *mut odb_exchange, handles: *const c_void) -> c_int;
        pub fn odb_exchange_allreduce(ex: *mut odb_exchange, dev_tile: *mut c_void, n_floats: u32, epilogue: c_int, cuda_stream: *mut c_void) -> c_int;
        pub fn odb_exchange_push(ex: *mut odb_exchange, dev_tile: *const c_void, n_floats: u32, cuda_stream: *mut c_void) -> c_int;
        pub fn odb_exchange_pull(ex: *mut odb_exchange, dev_tile: *mut c_void, n_floats: u32, epilogue: c_int, cuda_stream: *mut c_void) -> c_int;
        pub fn odb_scene_sample_exchange(scene: *mut odb_scene, ex: *mut odb_exchange, interval: f32, dev_out: *mut c_void, n_frames: u32, lag: c_int, epilogue: c_int, out_written: *mut c_int) -> c_int;
    }
}

// ---------------------------------------------------------------------------------------------------------------
/// The reference has no `Result` on this path (misuse panics, set.rs:160-163); neither has the shim: a failed call
/// panics with the library's message. `ODB_E_CUDA` at context creation means there is no CUDA device - the library
/// has no CPU fallback.
fn check(status: c_int) {
    if status != sys::ODB_OK {
        let msg = unsafe { CStr::from_ptr(sys::odb_last_error()) }.to_string_lossy().into_owned();
        panic!("oddio_b200 error {status}: {msg}");
    }
}

/// One CUDA device + stream + PCM arena. One per process and GPU; shared by every `Frames`, scene and mixer.
pub struct Context {
    raw: *mut sys::odb_ctx,
}
unsafe impl Send for Context {}
unsafe impl Sync for Context {}
impl Context {
    pub fn new(cuda_device: i32) -> Arc<Self> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::odb_ctx_create(cuda_device, &mut raw) });
        Arc::new(Self { raw })
    }
    /// Blocks until everything queued on the context's stream has finished.
    pub fn synchronize(&self) {
        check(unsafe { sys::odb_ctx_synchronize(self.raw) });
    }
    /// Page-locks an output buffer so that `sample` renders straight into it (no staging tile, no memcpy). The buffer
    /// must outlive the returned guard.
    pub fn pin<'a, T>(self: &Arc<Self>, buf: &'a mut [T]) -> Pinned<'a, T> {
        check(unsafe { sys::odb_pin_buffer(self.raw, buf.as_mut_ptr().cast(), std::mem::size_of_val(buf) as u64) });
        Pinned { ctx: self.clone(), buf }
    }
}
/// A caller buffer registered with the device for the guard's lifetime.
pub struct Pinned<'a, T> {
    ctx: Arc<Context>,
    pub buf: &'a mut [T],
}
impl<T> Drop for Pinned<'_, T> {
    fn drop(&mut self) {
        unsafe { sys::odb_unpin_buffer(self.ctx.raw, self.buf.as_mut_ptr().cast()) };
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::odb_ctx_destroy(self.raw) };
    }
}

// ---- Frame / Signal / Seek: the reference's traits, unchanged (frame.rs:4-13, signal.rs:14-58) -------------------
pub trait Frame {
    const ZERO: Self;
    fn channels(&self) -> &[Sample];
    fn channels_mut(&mut self) -> &mut [Sample];
}
impl Frame for Sample {
    const ZERO: Sample = 0.0;
    fn channels(&self) -> &[Sample] {
        std::slice::from_ref(self)
    }
    fn channels_mut(&mut self) -> &mut [Sample] {
        std::slice::from_mut(self)
    }
}
impl Frame for [Sample; 2] {
    const ZERO: [Sample; 2] = [0.0; 2];
    fn channels(&self) -> &[Sample] {
        self
    }
    fn channels_mut(&mut self) -> &mut [Sample] {
        self
    }
}
/// Frames the device path mixes: `Sample` (1 channel) and `[Sample; 2]`.
pub trait DeviceFrame: Frame + Copy + sealed::Sealed {
    const CHANNELS: c_int;
}
impl DeviceFrame for Sample {
    const CHANNELS: c_int = 1;
}
impl DeviceFrame for [Sample; 2] {
    const CHANNELS: c_int = 2;
}

pub trait Signal {
    type Frame;
    fn sample(&mut self, interval: f32, out: &mut [Self::Frame]);
    fn is_finished(&self) -> bool {
        false
    }
}
pub trait Seek: Signal {
    fn seek(&mut self, seconds: f32);
}
/// `oddio::run` (lib.rs:90-93).
pub fn run<S: Signal + ?Sized>(signal: &mut S, sample_rate: u32, out: &mut [S::Frame]) {
    let interval = 1.0 / sample_rate as f32;
    signal.sample(interval, out);
}

mod sealed {
    pub trait Sealed {}
    impl Sealed for f32 {}
    impl Sealed for [f32; 2] {}
}

// ---- Frames -----------------------------------------------------------------------------------------------------
/// `Arc<Frames<T>>` (frames.rs:16-77) resident in HBM.
pub struct Frames<T> {
    ctx: Arc<Context>,
    id: sys::odb_frames,
    rate: u32,
    len: usize,
    _t: PhantomData<T>,
}
unsafe impl<T> Send for Frames<T> {}
unsafe impl<T> Sync for Frames<T> {}
impl<T: DeviceFrame> Frames<T> {
    /// `Frames::from_slice` (frames.rs:26-47).
    pub fn from_slice(ctx: &Arc<Context>, rate: u32, samples: &[T]) -> Arc<Self> {
        let mut id = 0;
        check(unsafe { sys::odb_frames_from_slice(ctx.raw, rate, T::CHANNELS, samples.as_ptr().cast(), samples.len() as u64, &mut id) });
        Arc::new(Self { ctx: ctx.clone(), id, rate, len: samples.len(), _t: PhantomData })
    }
    /// `Frames::from_iter` (frames.rs:50-77).
    pub fn from_iter<I: IntoIterator<Item = T>>(ctx: &Arc<Context>, rate: u32, iter: I) -> Arc<Self> {
        let v: Vec<T> = iter.into_iter().collect();
        Self::from_slice(ctx, rate, &v)
    }
    /// Integer PCM as examples/wav.rs:30-46 decodes it, scaled on the device: `sample as f32 / (2^(bits-1) - 1) as f32`.
    pub fn from_i16(ctx: &Arc<Context>, rate: u32, samples: &[i16], bits_per_sample: u16) -> Arc<Self> {
        let n = samples.len() / T::CHANNELS as usize;
        let mut id = 0;
        check(unsafe { sys::odb_frames_from_i16(ctx.raw, rate, T::CHANNELS, samples.as_ptr(), n as u64, bits_per_sample as c_int, &mut id) });
        Arc::new(Self { ctx: ctx.clone(), id, rate, len: n, _t: PhantomData })
    }
    pub fn rate(&self) -> u32 {
        self.rate
    }
    pub fn len(&self) -> usize {
        self.len
    }
    pub fn is_empty(&self) -> bool {
        self.len == 0
    }
}
impl<T> Drop for Frames<T> {
    fn drop(&mut self) {
        unsafe { sys::odb_frames_release(self.ctx.raw, self.id) };
    }
}

// ---- the closed set of signal chains ------------------------------------------------------------------------------
/// Where a played signal lives: the owning scene / mixer and its source id. Shared (like the reference's `Arc` cells)
/// between the signal that was moved into `play` and the control handle the caller kept.
#[derive(Default)]
struct Binding {
    owner: std::sync::atomic::AtomicPtr<c_void>,
    src: std::sync::atomic::AtomicU64,
}
impl Binding {
    fn get(&self) -> Option<(*mut c_void, u64)> {
        let o = self.owner.load(std::sync::atomic::Ordering::Acquire);
        (!o.is_null()).then(|| (o, self.src.load(std::sync::atomic::Ordering::Acquire)))
    }
    fn set(&self, owner: *mut c_void, src: u64) {
        self.src.store(src, std::sync::atomic::Ordering::Release);
        self.owner.store(owner, std::sync::atomic::Ordering::Release);
    }
}

/// A signal chain the device path can play (sealed): `[Gain]([FixedGain]([Speed](FramesSignal | Cycle)))`.
pub trait DeviceSignal: chain::Sealed {
    type Frame: DeviceFrame;
    #[doc(hidden)]
    fn chain(&self) -> sys::odb_chain;
    #[doc(hidden)]
    fn bind(&mut self, owner: *mut c_void, src: u64);
    #[doc(hidden)]
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync>;
}
/// Chains that are `Seek` in the reference (frames.rs:209-214, cycle.rs:56-61, gain.rs:40-44): what
/// `SpatialSceneControl::play` accepts.
pub trait DeviceSeek: DeviceSignal {}
mod chain {
    pub trait Sealed {}
}

/// `FramesSignal<T>` (frames.rs:141-220). On the device once played; the host-side value only describes it.
pub struct FramesSignal<T: DeviceFrame> {
    data: Arc<Frames<T>>,
    start_seconds: f64,
    binding: Arc<Binding>,
}
/// `FramesSignalControl` (frames.rs:229-248).
pub struct FramesSignalControl {
    binding: Arc<Binding>,
}
impl<T: DeviceFrame + Send + Sync + 'static> FramesSignal<T> {
    /// `FramesSignal::new` (frames.rs:156-169).
    pub fn new(data: Arc<Frames<T>>, start_seconds: f64) -> (FramesSignalControl, Self) {
        let binding = Arc::new(Binding::default());
        (FramesSignalControl { binding: binding.clone() }, Self { data, start_seconds, binding })
    }
}
impl<T: DeviceFrame + Send + Sync + 'static> From<Arc<Frames<T>>> for FramesSignal<T> {
    fn from(data: Arc<Frames<T>>) -> Self {
        Self::new(data, 0.0).1
    }
}
impl FramesSignalControl {
    /// `FramesSignalControl::playback_position` (frames.rs:238-240). Panics before the signal is played.
    pub fn playback_position(&self) -> f64 {
        let (owner, src) = self.binding.get().expect("signal is not playing yet");
        let mut out = 0.0;
        check(unsafe { sys::odb_source_playback_position(owner, src, &mut out) });
        out
    }
    /// `FramesSignalControl::is_finished` (frames.rs:244-247).
    pub fn is_finished(&self) -> bool {
        let (owner, src) = self.binding.get().expect("signal is not playing yet");
        let mut out = 0;
        check(unsafe { sys::odb_source_frames_is_finished(owner, src, &mut out) });
        out != 0
    }
}
impl<T: DeviceFrame> chain::Sealed for FramesSignal<T> {}
impl<T: DeviceFrame + Send + Sync + 'static> DeviceSignal for FramesSignal<T> {
    type Frame = T;
    fn chain(&self) -> sys::odb_chain {
        sys::odb_chain { frames: self.data.id, start_seconds: self.start_seconds, flags: 0, speed: 1.0, fixed_gain_db: 0.0, gain_ratio: 1.0 }
    }
    fn bind(&mut self, owner: *mut c_void, src: u64) {
        self.binding.set(owner, src);
    }
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.data.clone()
    }
}
impl<T: DeviceFrame + Send + Sync + 'static> DeviceSeek for FramesSignal<T> {}

/// `Cycle<T>` (cycle.rs:6-61): loops a `Frames` block forever.
pub struct Cycle<T: DeviceFrame> {
    data: Arc<Frames<T>>,
    cursor: f64,
}
impl<T: DeviceFrame> Cycle<T> {
    /// `Cycle::new` (cycle.rs:15-20).
    pub fn new(data: Arc<Frames<T>>) -> Self {
        Self { data, cursor: 0.0 }
    }
    /// `Seek::seek` before the signal is played (cycle.rs:57-60): `(cursor + seconds * rate).rem_euclid(len)`.
    pub fn seek(&mut self, seconds: f32) {
        self.cursor = (self.cursor + f64::from(seconds) * f64::from(self.data.rate)).rem_euclid(self.data.len as f64);
    }
}
impl<T: DeviceFrame> chain::Sealed for Cycle<T> {}
impl<T: DeviceFrame + Send + Sync + 'static> DeviceSignal for Cycle<T> {
    type Frame = T;
    fn chain(&self) -> sys::odb_chain {
        sys::odb_chain { frames: self.data.id, start_seconds: self.cursor, flags: sys::ODB_CHAIN_CYCLE, speed: 1.0, fixed_gain_db: 0.0, gain_ratio: 1.0 }
    }
    fn bind(&mut self, _: *mut c_void, _: u64) {}
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.data.clone()
    }
}
impl<T: DeviceFrame + Send + Sync + 'static> DeviceSeek for Cycle<T> {}

/// `Speed<T>` (speed.rs:9-40). Not `Seek`, as in the reference.
pub struct Speed<S> {
    inner: S,
    speed: f32,
    binding: Arc<Binding>,
}
/// `SpeedControl` (speed.rs:43-55).
pub struct SpeedControl {
    binding: Arc<Binding>,
    speed: std::sync::atomic::AtomicU32,
}
impl<S: DeviceSignal> Speed<S> {
    /// `Speed::new` (speed.rs:16-23).
    pub fn new(inner: S) -> (SpeedControl, Self) {
        let binding = Arc::new(Binding::default());
        (SpeedControl { binding: binding.clone(), speed: std::sync::atomic::AtomicU32::new(1.0f32.to_bits()) }, Self { inner, speed: 1.0, binding })
    }
}
impl SpeedControl {
    /// `SpeedControl::speed` (speed.rs:47-49).
    pub fn speed(&self) -> f32 {
        f32::from_bits(self.speed.load(std::sync::atomic::Ordering::Relaxed))
    }
    /// `SpeedControl::set_speed` (speed.rs:52-54): takes effect at the next `sample`.
    pub fn set_speed(&mut self, factor: f32) {
        self.speed.store(factor.to_bits(), std::sync::atomic::Ordering::Relaxed);
        let (owner, src) = self.binding.get().expect("signal is not playing yet");
        check(unsafe { sys::odb_source_set_speed(owner, src, factor) });
    }
}
impl<S> chain::Sealed for Speed<S> {}
impl<S: DeviceSignal> DeviceSignal for Speed<S> {
    type Frame = S::Frame;
    fn chain(&self) -> sys::odb_chain {
        let mut c = self.inner.chain();
        c.flags |= sys::ODB_CHAIN_SPEED;
        c.speed = self.speed;
        c
    }
    fn bind(&mut self, owner: *mut c_void, src: u64) {
        self.binding.set(owner, src);
        self.inner.bind(owner, src);
    }
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.inner.keep_alive()
    }
}

/// `FixedGain<T>` (gain.rs:9-51). `Seek` when its inner signal is (gain.rs:40-44).
pub struct FixedGain<S> {
    inner: S,
    db: f32,
}
impl<S: DeviceSignal> FixedGain<S> {
    /// `FixedGain::new` (gain.rs:18-23): `gain = 10^(db / 20)`.
    pub fn new(inner: S, db: f32) -> Self {
        Self { inner, db }
    }
}
impl<S> chain::Sealed for FixedGain<S> {}
impl<S: DeviceSignal> DeviceSignal for FixedGain<S> {
    type Frame = S::Frame;
    fn chain(&self) -> sys::odb_chain {
        let mut c = self.inner.chain();
        c.flags |= sys::ODB_CHAIN_FIXED_GAIN;
        c.fixed_gain_db = self.db;
        c
    }
    fn bind(&mut self, owner: *mut c_void, src: u64) {
        self.inner.bind(owner, src);
    }
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.inner.keep_alive()
    }
}
impl<S: DeviceSeek> DeviceSeek for FixedGain<S> {}

/// `Gain<T>` (gain.rs:58-127): smoothed, dynamically adjustable gain. Not `Seek`.
pub struct Gain<S> {
    inner: S,
    ratio: f32,
    binding: Arc<Binding>,
}
/// `GainControl` (gain.rs:130-160).
pub struct GainControl {
    binding: Arc<Binding>,
}
impl<S: DeviceSignal> Gain<S> {
    /// `Gain::new` (gain.rs:66-93).
    pub fn new(inner: S) -> (GainControl, Self) {
        let binding = Arc::new(Binding::default());
        (GainControl { binding: binding.clone() }, Self { inner, ratio: 1.0, binding })
    }
    /// `Gain::set_gain` (gain.rs:81-83).
    pub fn set_gain(&mut self, db: f32) {
        self.set_amplitude_ratio(10.0f32.powf(db / 20.0));
    }
    /// `Gain::set_amplitude_ratio` (gain.rs:90-93): no smoothing before the signal is played.
    pub fn set_amplitude_ratio(&mut self, factor: f32) {
        self.ratio = factor;
    }
}
impl GainControl {
    /// `GainControl::set_gain` (gain.rs:143-145).
    pub fn set_gain(&mut self, db: f32) {
        let (owner, src) = self.binding.get().expect("signal is not playing yet");
        check(unsafe { sys::odb_source_set_gain_db(owner, src, db) });
    }
    /// `GainControl::set_amplitude_ratio` (gain.rs:157-159).
    pub fn set_amplitude_ratio(&mut self, factor: f32) {
        let (owner, src) = self.binding.get().expect("signal is not playing yet");
        check(unsafe { sys::odb_source_set_amplitude_ratio(owner, src, factor) });
    }
}
impl<S> chain::Sealed for Gain<S> {}
impl<S: DeviceSignal> DeviceSignal for Gain<S> {
    type Frame = S::Frame;
    fn chain(&self) -> sys::odb_chain {
        let mut c = self.inner.chain();
        c.flags |= sys::ODB_CHAIN_GAIN;
        c.gain_ratio = self.ratio;
        c
    }
    fn bind(&mut self, owner: *mut c_void, src: u64) {
        self.binding.set(owner, src);
        self.inner.bind(owner, src);
    }
    fn keep_alive(&self) -> Arc<dyn std::any::Any + Send + Sync> {
        self.inner.keep_alive()
    }
}

// ---- SpatialScene -----------------------------------------------------------------------------------------------
struct SceneRaw {
    raw: *mut sys::odb_scene,
    _ctx: Arc<Context>,
    keep: std::sync::Mutex<Vec<Arc<dyn std::any::Any + Send + Sync>>>, // Frames of playing signals (the library holds its own references too)
}
unsafe impl Send for SceneRaw {}
unsafe impl Sync for SceneRaw {}
impl Drop for SceneRaw {
    fn drop(&mut self) {
        unsafe { sys::odb_scene_destroy(self.raw) };
    }
}
/// The audio-thread half of a spatial scene (spatial.rs:160-189, :373-477).
pub struct SpatialScene {
    scene: Arc<SceneRaw>,
}
/// The control-thread half (spatial.rs:268-350).
pub struct SpatialSceneControl {
    scene: Arc<SceneRaw>,
}
/// `SpatialOptions` (spatial.rs:354-371).
#[derive(Clone, Copy, Debug)]
pub struct SpatialOptions {
    pub position: mint::Point3<f32>,
    pub velocity: mint::Vector3<f32>,
    pub radius: f32,
}
impl Default for SpatialOptions {
    fn default() -> Self {
        Self { position: [0.0; 3].into(), velocity: [0.0; 3].into(), radius: 0.1 }
    }
}
/// `Spatial` (spatial.rs:120-157): handle of one playing spatial signal.
pub struct Spatial {
    scene: Arc<SceneRaw>,
    src: sys::odb_source,
}
impl SpatialScene {
    /// `SpatialScene::new` (spatial.rs:170-188).
    pub fn new(ctx: &Arc<Context>) -> (SpatialSceneControl, Self) {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::odb_scene_create(ctx.raw, &mut raw) });
        let scene = Arc::new(SceneRaw { raw, _ctx: ctx.clone(), keep: Default::default() });
        (SpatialSceneControl { scene: scene.clone() }, Self { scene })
    }
}
impl SpatialSceneControl {
    /// `SpatialSceneControl::play` (spatial.rs:289-302).
    pub fn play<S: DeviceSeek<Frame = Sample>>(&mut self, mut signal: S, options: SpatialOptions) -> Spatial {
        let (p, v): ([f32; 3], [f32; 3]) = (options.position.into(), options.velocity.into());
        let mut src = 0;
        check(unsafe { sys::odb_scene_play(self.scene.raw, &signal.chain(), p.as_ptr(), v.as_ptr(), options.radius, &mut src) });
        signal.bind(self.scene.raw.cast(), src);
        self.scene.keep.lock().unwrap().push(signal.keep_alive());
        Spatial { scene: self.scene.clone(), src }
    }
    /// `SpatialSceneControl::play_buffered` (spatial.rs:314-340).
    pub fn play_buffered<S: DeviceSignal<Frame = Sample>>(&mut self, mut signal: S, options: SpatialOptions, max_distance: f32, rate: u32,
                                                          buffer_duration: f32) -> Spatial {
        let (p, v): ([f32; 3], [f32; 3]) = (options.position.into(), options.velocity.into());
        let mut src = 0;
        check(unsafe {
            sys::odb_scene_play_buffered(self.scene.raw, &signal.chain(), p.as_ptr(), v.as_ptr(), options.radius, max_distance, rate, buffer_duration, &mut src)
        });
        signal.bind(self.scene.raw.cast(), src);
        self.scene.keep.lock().unwrap().push(signal.keep_alive());
        Spatial { scene: self.scene.clone(), src }
    }
    /// `SpatialSceneControl::set_listener_rotation` (spatial.rs:345-349).
    pub fn set_listener_rotation(&mut self, rotation: mint::Quaternion<f32>) {
        let q = [rotation.v.x, rotation.v.y, rotation.v.z, rotation.s];
        check(unsafe { sys::odb_scene_set_listener_rotation(self.scene.raw, q.as_ptr()) });
    }
    /// `Spatial::set_motion` for many sources in one foreign call (no reference counterpart: one cheap Rust method call
    /// per source becomes one FFI call per batch).
    pub fn set_motion_many(&mut self, sources: &[&Spatial], positions: &[[f32; 3]], velocities: &[[f32; 3]], discontinuity: Option<&[u8]>) {
        assert!(positions.len() == sources.len() && velocities.len() == sources.len());
        let ids: Vec<u64> = sources.iter().map(|s| s.src).collect();
        check(unsafe {
            sys::odb_spatial_set_motion_many(self.scene.raw, ids.len() as u32, ids.as_ptr(), positions.as_ptr().cast(), velocities.as_ptr().cast(),
                                             discontinuity.map_or(ptr::null(), |d| d.as_ptr()))
        });
    }
}
impl Spatial {
    /// `Spatial::set_motion` (spatial.rs:137-149).
    pub fn set_motion(&mut self, position: mint::Point3<f32>, velocity: mint::Vector3<f32>, discontinuity: bool) {
        let (p, v): ([f32; 3], [f32; 3]) = (position.into(), velocity.into());
        check(unsafe { sys::odb_spatial_set_motion(self.scene.raw, self.src, p.as_ptr(), v.as_ptr(), discontinuity as c_int) });
    }
    /// `Spatial::is_finished` (spatial.rs:154-156).
    pub fn is_finished(&self) -> bool {
        let mut out = 0;
        check(unsafe { sys::odb_spatial_is_finished(self.scene.raw, self.src, &mut out) });
        out != 0
    }
}
impl Signal for SpatialScene {
    type Frame = [Sample; 2];
    /// `<SpatialScene as Signal>::sample` (spatial.rs:376-471). NOT wait-free: a CUDA launch and the wait for the
    /// 8 KiB tile (the deviation from signal.rs:11-13 is stated in DESIGN.md); it never allocates in steady state and
    /// never blocks on the control thread.
    fn sample(&mut self, interval: f32, out: &mut [[Sample; 2]]) {
        check(unsafe { sys::odb_scene_sample(self.scene.raw, interval, out.as_mut_ptr().cast(), out.len() as u32) });
    }
}

// ---- Mixer ------------------------------------------------------------------------------------------------------
struct MixerRaw {
    raw: *mut sys::odb_mixer,
    _ctx: Arc<Context>,
    keep: std::sync::Mutex<Vec<Arc<dyn std::any::Any + Send + Sync>>>,
}
unsafe impl Send for MixerRaw {}
unsafe impl Sync for MixerRaw {}
impl Drop for MixerRaw {
    fn drop(&mut self) {
        unsafe { sys::odb_mixer_destroy(self.raw) };
    }
}
/// `Mixer<T>` (mixer.rs:61-120): the audio-thread half.
pub struct Mixer<T> {
    mixer: Arc<MixerRaw>,
    _t: PhantomData<T>,
}
/// `MixerControl<T>` (mixer.rs:11-27).
pub struct MixerControl<T> {
    mixer: Arc<MixerRaw>,
    _t: PhantomData<T>,
}
/// `Mixed` (mixer.rs:30-44): handle of one playing signal.
pub struct Mixed {
    mixer: Arc<MixerRaw>,
    src: sys::odb_source,
}
unsafe impl<T> Send for Mixer<T> {}
unsafe impl<T> Send for MixerControl<T> {}
impl<T: DeviceFrame> Mixer<T> {
    /// `Mixer::new` (mixer.rs:70-81).
    pub fn new(ctx: &Arc<Context>) -> (MixerControl<T>, Self) {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::odb_mixer_create(ctx.raw, T::CHANNELS, &mut raw) });
        let mixer = Arc::new(MixerRaw { raw, _ctx: ctx.clone(), keep: Default::default() });
        (MixerControl { mixer: mixer.clone(), _t: PhantomData }, Self { mixer, _t: PhantomData })
    }
}
impl<T: DeviceFrame> MixerControl<T> {
    /// `MixerControl::play` (mixer.rs:18-26).
    pub fn play<S: DeviceSignal<Frame = T>>(&mut self, mut signal: S) -> Mixed {
        let mut src = 0;
        check(unsafe { sys::odb_mixer_play(self.mixer.raw, &signal.chain(), &mut src) });
        signal.bind(self.mixer.raw.cast(), src);
        self.mixer.keep.lock().unwrap().push(signal.keep_alive());
        Mixed { mixer: self.mixer.clone(), src }
    }
}
impl Mixed {
    /// `Mixed::stop` (mixer.rs:34-36).
    pub fn stop(&mut self) {
        check(unsafe { sys::odb_mixed_stop(self.mixer.raw, self.src) });
    }
    /// `Mixed::is_stopped` (mixer.rs:41-43).
    pub fn is_stopped(&self) -> bool {
        let mut out = 0;
        check(unsafe { sys::odb_mixed_is_stopped(self.mixer.raw, self.src, &mut out) });
        out != 0
    }
}
impl<T: DeviceFrame> Signal for Mixer<T> {
    type Frame = T;
    /// `<Mixer<T> as Signal>::sample` (mixer.rs:92-119).
    fn sample(&mut self, interval: f32, out: &mut [T]) {
        check(unsafe { sys::odb_mixer_sample(self.mixer.raw, interval, out.as_mut_ptr().cast(), out.len() as u32) });
    }
}

// ---- Tanh / Reinhard around an aggregator (tanh.rs:7-44, reinhard.rs:13-50) ----------------------------------------
/// Aggregators whose sum the device can post-process in its reduce phase.
pub trait DeviceAggregator: Signal + agg::Sealed {
    #[doc(hidden)]
    fn set_epilogue(&mut self, epilogue: c_int);
}
mod agg {
    pub trait Sealed {}
}
impl agg::Sealed for SpatialScene {}
impl DeviceAggregator for SpatialScene {
    fn set_epilogue(&mut self, e: c_int) {
        check(unsafe { sys::odb_scene_set_epilogue(self.scene.raw, e) });
    }
}
impl<T: DeviceFrame> agg::Sealed for Mixer<T> {}
impl<T: DeviceFrame> DeviceAggregator for Mixer<T> {
    fn set_epilogue(&mut self, e: c_int) {
        check(unsafe { sys::odb_mixer_set_epilogue(self.mixer.raw, e) });
    }
}
/// `Tanh<T>` (tanh.rs:7-44): `tanh` of every channel of the mixed output, applied in the kernel that finishes the sum.
pub struct Tanh<A>(A);
impl<A: DeviceAggregator> Tanh<A> {
    /// `Tanh::new` (tanh.rs:12-14).
    pub fn new(mut inner: A) -> Self {
        inner.set_epilogue(sys::ODB_EPILOGUE_TANH);
        Self(inner)
    }
}
impl<A: DeviceAggregator> Signal for Tanh<A> {
    type Frame = A::Frame;
    fn sample(&mut self, interval: f32, out: &mut [A::Frame]) {
        self.0.sample(interval, out)
    }
    fn is_finished(&self) -> bool {
        self.0.is_finished()
    }
}
/// `Reinhard<T>` (reinhard.rs:13-50): `x / (1 + |x|)`.
pub struct Reinhard<A>(A);
impl<A: DeviceAggregator> Reinhard<A> {
    /// `Reinhard::new` (reinhard.rs:18-20).
    pub fn new(mut inner: A) -> Self {
        inner.set_epilogue(sys::ODB_EPILOGUE_REINHARD);
        Self(inner)
    }
}
impl<A: DeviceAggregator> Signal for Reinhard<A> {
    type Frame = A::Frame;
    fn sample(&mut self, interval: f32, out: &mut [A::Frame]) {
        self.0.sample(interval, out)
    }
    fn is_finished(&self) -> bool {
        self.0.is_finished()
    }
}

// ---- offline render (examples/offline.rs:33-43) ------------------------------------------------------------------
impl SpatialScene {
    /// One callback quantised on the device as the example does before it writes the WAV file:
    /// `(sample * i16::MAX as f32) as i16`.
    pub fn sample_i16(&mut self, interval: f32, out: &mut [[i16; 2]]) {
        check(unsafe { sys::odb_scene_sample_i16(self.scene.raw, interval, out.as_mut_ptr().cast(), out.len() as u32) });
    }
    /// Sources in the seek set / the buffered set (`Set` derefs to a slice, set.rs:191-204).
    pub fn len(&self, buffered: bool) -> usize {
        let mut n = 0;
        check(unsafe { sys::odb_scene_len(self.scene.raw, buffered as c_int, &mut n) });
        n as usize
    }
}
impl<T: DeviceFrame> Mixer<T> {
    pub fn len(&self) -> usize {
        let mut n = 0;
        check(unsafe { sys::odb_mixer_len(self.mixer.raw, &mut n) });
        n as usize
    }
}

// ---- multi-GPU: one process per GPU, sources sharded, tiles summed over NVLink peer memory -------------------------
/// The per-box exchange of the mixed tile (no reference counterpart; include/oddio_b200.h "multi-GPU").
pub struct Exchange {
    raw: *mut sys::odb_exchange,
    _ctx: Arc<Context>,
}
unsafe impl Send for Exchange {}
impl Exchange {
    /// Local part: the inbox and its 64-byte handle. Gather the handles of all ranks (rank order) by any host-side
    /// means and pass them to [`Exchange::connect`].
    pub fn new(ctx: &Arc<Context>, rank: i32, world: i32, max_floats: u32, depth: i32) -> (Self, Vec<u8>) {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::odb_exchange_create(ctx.raw, rank, world, max_floats, depth, &mut raw) });
        let mut handle = vec![0u8; unsafe { sys::odb_exchange_handle_size() } as usize];
        check(unsafe { sys::odb_exchange_export(raw, handle.as_mut_ptr().cast()) });
        (Self { raw, _ctx: ctx.clone() }, handle)
    }
    pub fn connect(&mut self, handles_in_rank_order: &[u8]) {
        check(unsafe { sys::odb_exchange_connect(self.raw, handles_in_rank_order.as_ptr().cast()) });
    }
    /// In-place sum over the ranks of a device tile, then the epilogue.
    pub unsafe fn allreduce(&mut self, dev_tile: *mut c_void, n_floats: u32, epilogue: c_int) {
        check(sys::odb_exchange_allreduce(self.raw, dev_tile, n_floats, epilogue, ptr::null_mut()));
    }
    pub unsafe fn push(&mut self, dev_tile: *const c_void, n_floats: u32) {
        check(sys::odb_exchange_push(self.raw, dev_tile, n_floats, ptr::null_mut()));
    }
    pub unsafe fn pull(&mut self, dev_tile: *mut c_void, n_floats: u32, epilogue: c_int) {
        check(sys::odb_exchange_pull(self.raw, dev_tile, n_floats, epilogue, ptr::null_mut()));
    }
}
impl Drop for Exchange {
    fn drop(&mut self) {
        unsafe { sys::odb_exchange_destroy(self.raw) };
    }
}
impl SpatialScene {
    /// One callback of this rank's shard with the exchange folded into the callback kernel; returns true when
    /// `dev_out` (device memory, `2 * n_frames` f32) received the summed tile of callback `k - lag`.
    pub unsafe fn sample_exchange(&mut self, exchange: &mut Exchange, interval: f32, dev_out: *mut c_void, n_frames: u32, lag: i32, epilogue: c_int) -> bool {
        let mut written = 0;
        check(sys::odb_scene_sample_exchange(self.scene.raw, exchange.raw, interval, dev_out, n_frames, lag, epilogue, &mut written));
        written != 0
    }
    /// As `sample`, but the tile stays in device memory on the context's stream (no host synchronisation).
    pub unsafe fn sample_device(&mut self, interval: f32, dev_out: *mut c_void, n_frames: u32) {
        check(sys::odb_scene_sample_device(self.scene.raw, interval, dev_out, n_frames));
    }
}
