//! Raw declarations of include/rcb200.h (the subset the safe layer uses).
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct RcbCtx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RcbModel {
    _p: [u8; 0],
}
#[repr(C)]
pub struct RcbComm {
    _p: [u8; 0],
}
pub const RCB_UNIQUE_ID_BYTES: usize = 128;

/// `rcb_adaptive_params`: counts start at 1, the coded symbol gains `inc`, halving at `limit`.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RcbAdaptiveParams {
    pub k: u32,
    pub inc: u32,
    pub limit: u32,
}

/// `rcb_stream_state`: RangeCoder { lower_bound, range } (src/range_coder.rs:7-12 of the reference)
/// + Decoder::data (src/decoder.rs:8-12) + bookkeeping.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RcbStreamState {
    pub lower_bound: u64,
    pub range: u64,
    pub data: u64,
    pub consumed: u64,
    pub status: u32,
    pub pad: u32,
}

/// `rcb_restart_point`: the Encoder's RangeCoder in front of a symbol of a chunk (src/range_coder.rs:7-12,
/// range rounded down to a multiple of total_freq) and the bytes `Encoder::encode` has returned by then.
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RcbRestartPoint {
    pub lower_bound: u64,
    pub range: u64,
    pub code_bytes: u32,
    pub reserved: u32,
}

pub const RCB_OK: c_int = 0;
pub const RCB_ERR_ZERO_TOTAL: c_int = -3;
pub const RCB_ERR_ZERO_FREQ_SYMBOL: c_int = -4;
pub const RCB_ERR_LOWER_OVERFLOW: c_int = -5;
pub const RCB_ERR_UPPER_OVERFLOW: c_int = -6;
pub const RCB_ERR_SYMBOL_OUT_OF_RANGE: c_int = -7;
pub const RCB_ERR_OUT_CAPACITY: c_int = -8;
pub const RCB_ERR_TRUNCATED_STREAM: c_int = -9;
pub const RCB_ERR_NCCL: c_int = -13;
pub const RCB_ERR_RESTART_POINT: c_int = -14;

extern "C" {
    pub fn rcb_ctx_create(device: c_int, stream: *mut c_void, out: *mut *mut RcbCtx) -> c_int;
    pub fn rcb_ctx_destroy(ctx: *mut RcbCtx) -> c_int;
    pub fn rcb_strerror(code: c_int) -> *const c_char;

    pub fn rcb_model_create(ctx: *mut RcbCtx, k: u32, n_models: u64, out: *mut *mut RcbModel) -> c_int;
    pub fn rcb_model_destroy(m: *mut RcbModel) -> c_int;
    pub fn rcb_model_from_tables(
        ctx: *mut RcbCtx,
        m: *mut RcbModel,
        c: *const u32,
        cum: *const u32,
        total: *const u32,
    ) -> c_int;

    pub fn rcb_encode_bound(ctx: *mut RcbCtx, m: *const RcbModel, n: u64, sym_bytes: c_int, chunk: u64) -> u64;
    pub fn rcb_encode_host(
        ctx: *mut RcbCtx,
        syms: *const c_void,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        out: *mut u8,
        cap: u64,
        offsets: *mut u64,
        out_bytes: *mut u64,
    ) -> c_int;
    pub fn rcb_decode_host(
        ctx: *mut RcbCtx,
        stream: *const u8,
        offsets: *const u64,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        out: *mut c_void,
    ) -> c_int;

    // restart points: the Encoder's state every restart_syms symbols of a chunk (several decoder lanes per chunk)
    pub fn rcb_restart_points_per_chunk(chunk_syms: u64, restart_syms: u64) -> u64;
    pub fn rcb_encode_host_restart(
        ctx: *mut RcbCtx,
        syms: *const c_void,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        out: *mut u8,
        cap: u64,
        offsets: *mut u64,
        out_bytes: *mut u64,
        restart_syms: u64,
        restart: *mut RcbRestartPoint,
    ) -> c_int;
    pub fn rcb_decode_host_restart(
        ctx: *mut RcbCtx,
        stream: *const u8,
        offsets: *const u64,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        out: *mut c_void,
        restart_syms: u64,
        restart: *const RcbRestartPoint,
    ) -> c_int;

    pub fn rcb_stream_state_init(st: *mut RcbStreamState);
    pub fn rcb_encode_stream(
        ctx: *mut RcbCtx,
        st: *mut RcbStreamState,
        h_syms: *const c_void,
        n_syms: u64,
        sym_bytes: c_int,
        m: *const RcbModel,
        h_out: *mut u8,
        out_cap: u64,
        h_n_out: *mut u64,
        h_per_symbol: *mut u32,
        finish: c_int,
    ) -> c_int;
    pub fn rcb_decode_stream(
        ctx: *mut RcbCtx,
        st: *mut RcbStreamState,
        h_code: *const u8,
        code_len: u64,
        n_syms: u64,
        sym_bytes: c_int,
        m: *const RcbModel,
        h_syms_out: *mut c_void,
    ) -> c_int;

    // device memory + the device-pointer entry points (sharding over several GPUs)
    pub fn rcb_device_count() -> c_int;
    pub fn rcb_device_alloc(ctx: *mut RcbCtx, bytes: u64, d_out: *mut *mut c_void) -> c_int;
    pub fn rcb_device_free(ctx: *mut RcbCtx, d_ptr: *mut c_void) -> c_int;
    pub fn rcb_copy_to_device(ctx: *mut RcbCtx, d_dst: *mut c_void, h_src: *const c_void, bytes: u64) -> c_int;
    pub fn rcb_copy_to_host(ctx: *mut RcbCtx, h_dst: *mut c_void, d_src: *const c_void, bytes: u64) -> c_int;
    pub fn rcb_histogram(
        ctx: *mut RcbCtx,
        d_syms: *const c_void,
        n_syms: u64,
        sym_bytes: c_int,
        k: u32,
        chunk_syms: u64,
        d_counts: *mut c_void,
    ) -> c_int;
    pub fn rcb_model_from_counts(ctx: *mut RcbCtx, m: *mut RcbModel, d_counts: *const c_void, count_bytes: c_int) -> c_int;
    pub fn rcb_encode_chunks(
        ctx: *mut RcbCtx,
        d_syms: *const c_void,
        n_syms: u64,
        sym_bytes: c_int,
        chunk_syms: u64,
        m: *const RcbModel,
        d_out: *mut u8,
        out_cap: u64,
        d_offsets: *mut u64,
        d_status: *mut u32,
        h_out_bytes: *mut u64,
    ) -> c_int;

    // adaptive-per-symbol table: device buffers
    pub fn rcb_adaptive_encode_bound(p: *const RcbAdaptiveParams, n_syms: u64, chunk_syms: u64) -> u64;
    pub fn rcb_adaptive_encode_chunks(
        ctx: *mut RcbCtx,
        d_syms: *const c_void,
        n_syms: u64,
        sym_bytes: c_int,
        chunk_syms: u64,
        p: *const RcbAdaptiveParams,
        d_out: *mut u8,
        out_cap: u64,
        d_offsets: *mut u64,
        d_status: *mut u32,
        h_out_bytes: *mut u64,
    ) -> c_int;
    pub fn rcb_adaptive_decode_chunks(
        ctx: *mut RcbCtx,
        d_stream: *const u8,
        d_offsets: *const u64,
        n_syms: u64,
        sym_bytes: c_int,
        chunk_syms: u64,
        p: *const RcbAdaptiveParams,
        d_syms_out: *mut c_void,
        d_status: *mut u32,
    ) -> c_int;

    // the path's only exchange step: one NCCL all-reduce of the K u64 counts (include/rcb200.h)
    pub fn rcb_comm_unique_id(id: *mut u8) -> c_int;
    pub fn rcb_comm_init_rank(ctx: *mut RcbCtx, id: *const u8, n_ranks: c_int, rank: c_int, out: *mut *mut RcbComm) -> c_int;
    pub fn rcb_comm_init_all(ctxs: *const *mut RcbCtx, n_ctx: c_int, out: *mut *mut RcbComm) -> c_int;
    pub fn rcb_comm_destroy(comm: *mut RcbComm) -> c_int;
    pub fn rcb_comm_last_error(comm: *const RcbComm) -> *const c_char;
    pub fn rcb_allreduce_counts(ctx: *mut RcbCtx, comm: *mut RcbComm, d_counts: *mut c_void, k: u32) -> c_int;
    pub fn rcb_allreduce_counts_multi(
        ctxs: *const *mut RcbCtx,
        comms: *const *mut RcbComm,
        d_counts: *const *mut c_void,
        k: u32,
        n_ctx: c_int,
    ) -> c_int;

    pub fn rcb_frame_bound(k: u32, n_chunks: u64, per_chunk: c_int, payload: u64) -> u64;
    pub fn rcb_frame_encode_host(
        ctx: *mut RcbCtx,
        syms: *const c_void,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        frame: *mut u8,
        cap: u64,
        frame_bytes: *mut u64,
    ) -> c_int;
    pub fn rcb_frame_bound_restart(
        k: u32,
        n_chunks: u64,
        per_chunk: c_int,
        payload: u64,
        chunk_syms: u64,
        restart_syms: u64,
    ) -> u64;
    pub fn rcb_frame_encode_host_restart(
        ctx: *mut RcbCtx,
        syms: *const c_void,
        n: u64,
        sym_bytes: c_int,
        chunk: u64,
        m: *const RcbModel,
        restart_syms: u64,
        frame: *mut u8,
        cap: u64,
        frame_bytes: *mut u64,
    ) -> c_int;
    pub fn rcb_frame_decode_host(
        ctx: *mut RcbCtx,
        frame: *const u8,
        len: u64,
        out: *mut c_void,
        out_cap_bytes: u64,
        n_syms: *mut u64,
    ) -> c_int;
}
