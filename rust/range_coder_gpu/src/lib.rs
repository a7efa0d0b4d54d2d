//! `range_coder_gpu`: the reference crate's hot loops on a B200, behind the reference's own types.
//!
//! A caller of `range_coder` writes (examples/sample_impl.rs:92-98, 113-120 of the reference)
//! ```ignore
//! let mut encoder = Encoder::new();
//! for &i in &data { encoder.encode(&sd, i); }
//! let code = encoder.finish();
//! ```
//! and with this crate
//! ```ignore
//! let gpu = Gpu::new(0);
//! let table = gpu.snapshot(&sd, sd.alphabet_count());
//! let enc = gpu.encode_chunks(&table, &data, data.len() as u64);   // enc.chunk(0) == code
//! let back = gpu.decode_chunks(&table, &enc, data.len() as u64);
//! ```
//! Every chunk's bytes are exactly what `Encoder::finish` returns for that chunk's symbols.
//! There is no CPU fallback: without a CUDA device `Gpu::new` panics with the library's message.
pub mod ffi;

use range_coder::error::RangeCoderError;
use range_coder::PModel;
use std::ffi::CStr;
use std::os::raw::{c_int, c_void};

/// One library context (one CUDA device + stream).  Not `Sync`: one per GPU per thread, like the
/// reference's single-threaded owned values.
pub struct Gpu {
    ctx: *mut ffi::RcbCtx,
}

/// Dense device-side snapshot of a `PModel` (src/pmodel.rs:4-13): `c_freq`, `cum_freq`, `total_freq`.
pub struct Table<'g> {
    gpu: &'g Gpu,
    m: *mut ffi::RcbModel,
    k: usize,
}

/// Concatenated per-chunk code streams; chunk `i` is `stream[offsets[i]..offsets[i+1]]`.
pub struct Encoded {
    pub stream: Vec<u8>,
    pub offsets: Vec<u64>,
    pub chunk_syms: u64,
    /// Restart points (include/rcb200.h): the Encoder's state every `restart_syms` symbols of a chunk.  Side
    /// information for `decode_chunks` (several GPU lanes per chunk); `stream` is unchanged by it.
    pub restart_syms: u64,
    pub restart: Vec<ffi::RcbRestartPoint>,
}

/// Restart points: 16 parts per chunk for chunks of >= 32 Ki symbols, else 4, when the part is a whole number
/// of 64-symbol units (0: none).
pub fn default_restart_syms(chunk_syms: u64) -> u64 {
    if chunk_syms >= 32768 && chunk_syms % 1024 == 0 {
        chunk_syms / 16
    } else if chunk_syms % 256 == 0 {
        chunk_syms / 4
    } else {
        0
    }
}

impl Encoded {
    pub fn chunk(&self, i: usize) -> &[u8] {
        &self.stream[self.offsets[i] as usize..self.offsets[i + 1] as usize]
    }
}

fn message(rc: c_int) -> String {
    unsafe { CStr::from_ptr(ffi::rcb_strerror(rc)).to_string_lossy().into_owned() }
}

/// Mirror of the reference's failure behaviour: its encoder/decoder `unwrap()` a `RangeCoderError`
/// (src/encoder.rs:33, src/decoder.rs:33,49) or hit a panic (divide by zero, `pop_front().unwrap()`),
/// so the bulk calls panic with the same error type; `try_*` variants return it instead.
fn to_result(rc: c_int) -> Result<(), GpuError> {
    match rc {
        ffi::RCB_OK => Ok(()),
        ffi::RCB_ERR_LOWER_OVERFLOW => Err(GpuError::RangeCoder(RangeCoderError::LowerBoundOverflow {
            lower_bound: 0,
            add_val: 0,
            range: 0,
        })),
        ffi::RCB_ERR_UPPER_OVERFLOW => {
            Err(GpuError::RangeCoder(RangeCoderError::UpperBoundOverflow { lower_bound: 0, range: 0 }))
        }
        e => Err(GpuError::Library(e, message(e))),
    }
}

#[derive(Debug)]
pub enum GpuError {
    RangeCoder(RangeCoderError),
    Library(c_int, String),
}

impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            GpuError::RangeCoder(e) => write!(f, "{}", e),
            GpuError::Library(c, m) => write!(f, "rcb200 error {}: {}", c, m),
        }
    }
}
impl std::error::Error for GpuError {}

impl Gpu {
    pub fn new(device: i32) -> Self {
        Self::try_new(device).unwrap()
    }
    pub fn try_new(device: i32) -> Result<Self, GpuError> {
        let mut ctx = std::ptr::null_mut();
        to_result(unsafe { ffi::rcb_ctx_create(device as c_int, std::ptr::null_mut(), &mut ctx) })?;
        Ok(Gpu { ctx })
    }

    /// Read `k` entries of any `PModel` into a device table.
    pub fn snapshot<T: PModel>(&self, pmodel: &T, k: usize) -> Table<'_> {
        let c: Vec<u32> = (0..k).map(|i| pmodel.c_freq(i)).collect();
        let cum: Vec<u32> = (0..k).map(|i| pmodel.cum_freq(i)).collect();
        let total = pmodel.total_freq();
        let mut m = std::ptr::null_mut();
        unsafe {
            to_result(ffi::rcb_model_create(self.ctx, k as u32, 1, &mut m)).unwrap();
            to_result(ffi::rcb_model_from_tables(self.ctx, m, c.as_ptr(), cum.as_ptr(), &total)).unwrap();
        }
        Table { gpu: self, m, k }
    }

    /// One independent `Encoder` run (new .. encode* .. finish) per `chunk_syms` symbols.
    pub fn try_encode_chunks(&self, t: &Table<'_>, symbols: &[u8], chunk_syms: u64) -> Result<Encoded, GpuError> {
        let n = symbols.len() as u64;
        let n_chunks = (n + chunk_syms - 1) / chunk_syms;
        unsafe {
            let cap = ffi::rcb_encode_bound(self.ctx, t.m, n, 1, chunk_syms) + 16;
            let mut stream = vec![0u8; cap as usize];
            let mut offsets = vec![0u64; n_chunks as usize + 1];
            let mut bytes = 0u64;
            let mut restart_syms = default_restart_syms(chunk_syms);
            let per = ffi::rcb_restart_points_per_chunk(chunk_syms, restart_syms);
            if per == 0 {
                restart_syms = 0;
            }
            let mut restart = vec![ffi::RcbRestartPoint::default(); (n_chunks * per) as usize];
            to_result(ffi::rcb_encode_host_restart(
                self.ctx,
                symbols.as_ptr() as *const c_void,
                n,
                1,
                chunk_syms,
                t.m,
                stream.as_mut_ptr(),
                cap,
                offsets.as_mut_ptr(),
                &mut bytes,
                restart_syms,
                if restart.is_empty() { std::ptr::null_mut() } else { restart.as_mut_ptr() },
            ))?;
            stream.truncate(bytes as usize);
            Ok(Encoded { stream, offsets, chunk_syms, restart_syms, restart })
        }
    }
    pub fn encode_chunks(&self, t: &Table<'_>, symbols: &[u8], chunk_syms: u64) -> Encoded {
        self.try_encode_chunks(t, symbols, chunk_syms).unwrap()
    }

    /// One `Decoder::new` + `n` x `Decoder::decode` per chunk; the lookup closure of the reference
    /// (`find_index`, examples/sample_impl.rs:27-45: smallest index whose next cum_freq exceeds the
    /// decoded frequency, clamped to the last symbol) is built in.
    pub fn try_decode_chunks(&self, t: &Table<'_>, enc: &Encoded, n_syms: u64) -> Result<Vec<u8>, GpuError> {
        let mut padded = enc.stream.clone();
        padded.resize((enc.stream.len() + 15) / 16 * 16 + 16, 0); // device reader fetches 16-byte pieces
        let mut out = vec![0u8; n_syms as usize];
        to_result(unsafe {
            ffi::rcb_decode_host_restart(
                self.ctx,
                padded.as_ptr(),
                enc.offsets.as_ptr(),
                n_syms,
                1,
                enc.chunk_syms,
                t.m,
                out.as_mut_ptr() as *mut c_void,
                enc.restart_syms,
                if enc.restart.is_empty() { std::ptr::null() } else { enc.restart.as_ptr() },
            )
        })?;
        Ok(out)
    }
    pub fn decode_chunks(&self, t: &Table<'_>, enc: &Encoded, n_syms: u64) -> Vec<u8> {
        self.try_decode_chunks(t, enc, n_syms).unwrap()
    }

    /// Self-describing container (header + table + offsets + payload), see include/rcb200.h.
    pub fn encode_frame(&self, t: &Table<'_>, symbols: &[u8], chunk_syms: u64) -> Result<Vec<u8>, GpuError> {
        let n = symbols.len() as u64;
        let n_chunks = (n + chunk_syms - 1) / chunk_syms;
        unsafe {
            let payload = ffi::rcb_encode_bound(self.ctx, t.m, n, 1, chunk_syms);
            // version-2 frame: restart points next to the offsets (several decoder lanes per chunk)
            let restart_syms = default_restart_syms(chunk_syms);
            let cap = ffi::rcb_frame_bound_restart(t.k as u32, n_chunks, 0, payload, chunk_syms, restart_syms);
            let mut frame = vec![0u8; cap as usize];
            let mut bytes = 0u64;
            to_result(ffi::rcb_frame_encode_host_restart(
                self.ctx,
                symbols.as_ptr() as *const c_void,
                n,
                1,
                chunk_syms,
                t.m,
                restart_syms,
                frame.as_mut_ptr(),
                cap,
                &mut bytes,
            ))?;
            frame.truncate(bytes as usize);
            Ok(frame)
        }
    }
    pub fn decode_frame(&self, frame: &[u8]) -> Result<Vec<u8>, GpuError> {
        let mut n = 0u64;
        // first call sizes the output (returns OUT_CAPACITY with n set), second decodes
        let rc = unsafe {
            ffi::rcb_frame_decode_host(self.ctx, frame.as_ptr(), frame.len() as u64, std::ptr::null_mut(), 0, &mut n)
        };
        if rc != ffi::RCB_OK && rc != ffi::RCB_ERR_OUT_CAPACITY {
            to_result(rc)?;
        }
        let mut out = vec![0u8; n as usize];
        to_result(unsafe {
            ffi::rcb_frame_decode_host(
                self.ctx,
                frame.as_ptr(),
                frame.len() as u64,
                out.as_mut_ptr() as *mut c_void,
                n,
                &mut n,
            )
        })?;
        Ok(out)
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe {
            ffi::rcb_ctx_destroy(self.ctx);
        }
    }
}
impl<'g> Drop for Table<'g> {
    fn drop(&mut self) {
        let _ = self.gpu; // tables must not outlive their context: enforced by the borrow
        unsafe {
            ffi::rcb_model_destroy(self.m);
        }
    }
}

/// Drop-in for `range_coder::Encoder` with the per-symbol API (src/encoder.rs:14-46): the coder state
/// lives in `st` between calls and each `encode` runs on the GPU.  For throughput use
/// `Gpu::encode_chunks`; this exists so code written against the reference compiles unchanged.
pub struct Encoder<'g> {
    gpu: &'g Gpu,
    st: ffi::RcbStreamState,
    data: Vec<u8>,
}

impl<'g> Encoder<'g> {
    pub fn new(gpu: &'g Gpu) -> Self {
        let mut st = ffi::RcbStreamState::default();
        unsafe { ffi::rcb_stream_state_init(&mut st) };
        Encoder { gpu, st, data: Vec::new() }
    }
    /// Returns the number of bytes this symbol produced, like `Encoder::encode`.
    pub fn encode(&mut self, t: &Table<'_>, index: usize) -> u32 {
        let sym = [index as u16];
        let mut buf = [0u8; 16];
        let mut n = 0u64;
        let mut per = 0u32;
        to_result(unsafe {
            ffi::rcb_encode_stream(
                self.gpu.ctx,
                &mut self.st,
                sym.as_ptr() as *const c_void,
                1,
                2,
                t.m,
                buf.as_mut_ptr(),
                16,
                &mut n,
                &mut per,
                0,
            )
        })
        .unwrap();
        self.data.extend_from_slice(&buf[..n as usize]);
        per
    }
    pub fn finish(mut self) -> Vec<u8> {
        let mut buf = [0u8; 16];
        let mut n = 0u64;
        to_result(unsafe {
            ffi::rcb_encode_stream(
                self.gpu.ctx,
                &mut self.st,
                std::ptr::null(),
                0,
                1,
                std::ptr::null(),
                buf.as_mut_ptr(),
                16,
                &mut n,
                std::ptr::null_mut(),
                1,
            )
        })
        .unwrap();
        self.data.extend_from_slice(&buf[..n as usize]);
        self.data
    }
}

/// Drop-in for `range_coder::Decoder` (src/decoder.rs:14-54) over `rcb_decode_stream`.
pub struct Decoder<'g> {
    gpu: &'g Gpu,
    st: ffi::RcbStreamState,
    code: Vec<u8>,
}

impl<'g> Decoder<'g> {
    pub fn new(gpu: &'g Gpu, code: Vec<u8>) -> Self {
        let mut st = ffi::RcbStreamState::default();
        unsafe { ffi::rcb_stream_state_init(&mut st) };
        Decoder { gpu, st, code }
    }
    pub fn decode(&mut self, t: &Table<'_>) -> usize {
        let mut sym = [0u16];
        to_result(unsafe {
            ffi::rcb_decode_stream(
                self.gpu.ctx,
                &mut self.st,
                self.code.as_ptr(),
                self.code.len() as u64,
                1,
                2,
                t.m,
                sym.as_mut_ptr() as *mut c_void,
            )
        })
        .unwrap();
        sym[0] as usize
    }
}

/// A static table shared by chunks that live on several GPUs, one host thread driving all of them
/// (`rcb_comm_init_all`).  GPU `g` owns the chunks `[n_chunks*g/G, n_chunks*(g+1)/G)`; the table is
/// what the reference's caller builds with `add_alphabet_freq` over ALL the data + `calc_cum`
/// (examples/sample_impl.rs:77-81): per-GPU histograms, ONE all-reduce of the K counts over NVLink,
/// the same deterministic table on every GPU.  No payload byte crosses GPUs.
pub struct MultiGpu {
    gpus: Vec<Gpu>,
    comms: Vec<*mut ffi::RcbComm>,
}

impl MultiGpu {
    pub fn new(n_gpus: usize) -> Result<Self, GpuError> {
        let gpus = (0..n_gpus).map(|g| Gpu::try_new(g as i32)).collect::<Result<Vec<_>, _>>()?;
        let ctxs: Vec<*mut ffi::RcbCtx> = gpus.iter().map(|g| g.ctx).collect();
        let mut comms = vec![std::ptr::null_mut(); n_gpus];
        to_result(unsafe { ffi::rcb_comm_init_all(ctxs.as_ptr(), n_gpus as c_int, comms.as_mut_ptr()) })?;
        Ok(MultiGpu { gpus, comms })
    }

    /// histogram (sharded) -> all-reduce -> table -> encode; the per-GPU streams concatenate in chunk order.
    pub fn encode_chunks(&self, symbols: &[u8], k: u32, chunk_syms: u64) -> Result<Encoded, GpuError> {
        let g_n = self.gpus.len() as u64;
        let n = symbols.len() as u64;
        let n_chunks = (n + chunk_syms - 1) / chunk_syms;
        let first: Vec<u64> = (0..=g_n).map(|g| n.min(n_chunks * g / g_n * chunk_syms)).collect();
        let mut d_syms = vec![std::ptr::null_mut::<c_void>(); g_n as usize];
        let mut d_counts = vec![std::ptr::null_mut::<c_void>(); g_n as usize];
        unsafe {
            for (g, gpu) in self.gpus.iter().enumerate() {
                let cnt = first[g + 1] - first[g];
                to_result(ffi::rcb_device_alloc(gpu.ctx, cnt, &mut d_syms[g]))?;
                to_result(ffi::rcb_device_alloc(gpu.ctx, k as u64 * 8, &mut d_counts[g]))?;
                let src = symbols[first[g] as usize..].as_ptr() as *const c_void;
                to_result(ffi::rcb_copy_to_device(gpu.ctx, d_syms[g], src, cnt))?;
                to_result(ffi::rcb_histogram(gpu.ctx, d_syms[g], cnt, 1, k, 0, d_counts[g]))?;
            }
            let ctxs: Vec<*mut ffi::RcbCtx> = self.gpus.iter().map(|g| g.ctx).collect();
            to_result(ffi::rcb_allreduce_counts_multi(
                ctxs.as_ptr(),
                self.comms.as_ptr(),
                d_counts.as_ptr(),
                k,
                g_n as c_int,
            ))?;
            let mut enc = Encoded { stream: Vec::new(), offsets: vec![0u64], chunk_syms, restart_syms: 0, restart: Vec::new() };
            for (g, gpu) in self.gpus.iter().enumerate() {
                let cnt = first[g + 1] - first[g];
                let chunks = (cnt + chunk_syms - 1) / chunk_syms;
                let mut m = std::ptr::null_mut();
                to_result(ffi::rcb_model_create(gpu.ctx, k, 1, &mut m))?;
                to_result(ffi::rcb_model_from_counts(gpu.ctx, m, d_counts[g], 8))?;
                let cap = ffi::rcb_encode_bound(gpu.ctx, m, cnt, 1, chunk_syms) + 16;
                let (mut d_out, mut d_off) = (std::ptr::null_mut::<c_void>(), std::ptr::null_mut::<c_void>());
                to_result(ffi::rcb_device_alloc(gpu.ctx, cap, &mut d_out))?;
                to_result(ffi::rcb_device_alloc(gpu.ctx, (chunks + 1) * 8, &mut d_off))?;
                let mut bytes = 0u64;
                to_result(ffi::rcb_encode_chunks(
                    gpu.ctx,
                    d_syms[g],
                    cnt,
                    1,
                    chunk_syms,
                    m,
                    d_out as *mut u8,
                    cap,
                    d_off as *mut u64,
                    std::ptr::null_mut(),
                    &mut bytes,
                ))?;
                let base = enc.stream.len();
                let k0 = enc.offsets.len() - 1;
                enc.stream.resize(base + bytes as usize, 0);
                enc.offsets.resize(k0 + chunks as usize + 1, 0);
                to_result(ffi::rcb_copy_to_host(gpu.ctx, enc.stream[base..].as_mut_ptr() as *mut c_void, d_out, bytes))?;
                to_result(ffi::rcb_copy_to_host(
                    gpu.ctx,
                    enc.offsets[k0..].as_mut_ptr() as *mut c_void,
                    d_off,
                    (chunks + 1) * 8,
                ))?;
                for o in &mut enc.offsets[k0..] {
                    *o += base as u64; // local -> global offsets
                }
                ffi::rcb_device_free(gpu.ctx, d_out);
                ffi::rcb_device_free(gpu.ctx, d_off);
                ffi::rcb_device_free(gpu.ctx, d_syms[g]);
                ffi::rcb_device_free(gpu.ctx, d_counts[g]);
                ffi::rcb_model_destroy(m);
            }
            Ok(enc)
        }
    }
}

impl Drop for MultiGpu {
    fn drop(&mut self) {
        for &k in &self.comms {
            unsafe {
                ffi::rcb_comm_destroy(k);
            }
        }
    }
}
