//! Rust shim over the C ABI in include/fqgrep_b200.h, for the reference's own crate to link.
//!
//! Mirrors, name for name, the seam the reference exposes for the read-matching path:
//!   * `MatcherOpts`                                   fqgrep src/lib/matcher.rs:20-28
//!   * `trait Matcher` (`opts`, `bases_match`, `read_match`)          :150-193
//!   * `MatcherFactory::{new_matcher, new_query_name_matcher}`        :641-686 (five / two arguments, same order)
//!   * `validate_fixed_pattern`                                        :128-147
//! and adds what a GPU needs that a per-record virtual call cannot give: `BatchMatcher::match_batch` for the batch seam
//! (src/lib/seq_io.rs:314-326) and `FastqStream`, the ring of batches in flight that replaces reader thread -> worker
//! pool -> ordered main thread (src/main.rs:61-87, src/lib/seq_io.rs:149-198, 253-282, 327-364).
//!
//! NOTE: shipped as source only.  The build image has no cargo/rustc, so this file is compile-unverified; every
//! declaration below is checked against the C header by tests/test_compile_cpu.py (names) and kept next to it by hand
//! (signatures).
use ahash::AHashSet;
use anyhow::{bail, Result};
use seq_io::fastq::{Record, RefRecord};
use std::ffi::{c_char, c_int, c_void, CStr, CString};

/// Common options for pattern matchers (matcher.rs:20-28).
#[derive(Copy, Clone, Debug)]
pub struct MatcherOpts {
    pub invert_match: bool,
    pub reverse_complement: bool,
}

#[repr(i32)]
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Kind { Fixed = 0, FixedSet = 1, Regex = 2, RegexSet = 3, Names = 4, BitMask = 5, BitMaskSet = 6 }

#[repr(i32)]
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Mode { Single = 0, Interleaved = 1, Paired = 2 }

/// fqg_result
#[repr(C)]
#[derive(Default, Debug, Clone)]
pub struct FqgResult {
    pub n_records: u64, pub n_units: u64, pub n_selected: u64,
    pub h2d_ms: f64, pub kernel_ms: f64, pub d2h_ms: f64,
    pub h2d_bytes: u64, pub d2h_bytes: u64,
}
/// fqg_dfa_info
#[repr(C)]
#[derive(Default, Debug)]
pub struct FqgDfaInfo {
    pub n_states: u32, pub n_classes: u32, pub start: u32, pub accept_from: u32, pub tier: u32, pub table_bytes: u32,
}
/// fqg_span: one selected record in write-back order
#[repr(C)]
#[derive(Default, Debug, Copy, Clone)]
pub struct FqgSpan { pub offset: u64, pub len: u32, pub flags: u32 }
pub const FQG_SPAN_R2: u32 = 1;
pub const FQG_SPAN_NORMALISE: u32 = 2;
/// fqg_stream_opts
#[repr(C)]
#[derive(Default, Debug, Copy, Clone)]
pub struct FqgStreamOpts { pub n_slots: u32, pub want_spans: u32, pub slot_bytes: usize }
/// fqg_batch
#[repr(C)]
#[derive(Debug, Copy, Clone)]
pub struct FqgBatch {
    pub seq: u64, pub n_records: u64, pub n_units: u64, pub n_selected: u64,
    pub unit_mask: *const u32, pub spans: *const FqgSpan, pub n_spans: u64,
    pub r1: *const u8, pub r2: *const u8, pub n1: usize, pub n2: usize,
    pub h2d_ms: f64, pub kernel_ms: f64,
}

#[link(name = "fqgrep_b200")]
extern "C" {
    pub fn fqg_last_error() -> *const c_char;
    pub fn fqg_version() -> *const c_char;
    pub fn fqg_matcher_new(kind: c_int, blob: *const u8, offsets: *const u64, n: u32, opts: u32, out: *mut *mut c_void) -> c_int;
    pub fn fqg_matcher_free(m: *mut c_void);
    pub fn fqg_validate_fixed_pattern(pattern: *const c_char, protein: c_int) -> c_int;
    pub fn fqg_iupac_expand(pattern: *const c_char, to_regex: c_int, out: *mut c_char, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fqg_matcher_dfa_info(m: *const c_void, out: *mut FqgDfaInfo) -> c_int;
    pub fn fqg_matcher_dfa_export(m: *const c_void, byte_class: *mut u8, next: *mut u32) -> c_int;
    pub fn fqg_ctx_create(device: c_int, out: *mut *mut c_void) -> c_int;
    pub fn fqg_ctx_destroy(ctx: *mut c_void);
    pub fn fqg_device_count() -> c_int;
    pub fn fqg_alloc_pinned(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn fqg_free_pinned(p: *mut c_void);
    pub fn fqg_match_bases_device(ctx: *mut c_void, m: *const c_void, d_bases: *const u8, d_start: *const u32, d_end: *const u32, n_reads: u64,
                                  d_mask_out: *mut u32, d_or_mask: *const u32, d_count: *mut u64, avg_read_len: u32, stream: *mut c_void) -> c_int;
    pub fn fqg_match_bases_host(ctx: *mut c_void, m: *const c_void, bases: *const u8, offsets: *const u64, n_reads: u64,
                                mask_out: *mut u32, n_selected: *mut u64) -> c_int;
    pub fn fqg_grep_fastq_host(ctx: *mut c_void, m: *const c_void, mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize,
                               unit_mask_out: *mut u32, unit_mask_words: usize, res: *mut FqgResult) -> c_int;
    pub fn fqg_grep_fastq_host_write(ctx: *mut c_void, m: *const c_void, mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize,
                                     out: *mut u8, out_cap: usize, out_len: *mut usize, out2: *mut u8, out2_cap: usize, out2_len: *mut usize,
                                     res: *mut FqgResult) -> c_int;
    pub fn fqg_grep_fastq_host_multi(ctxs: *const *mut c_void, n_ctx: u32, m: *const c_void, mode: c_int, r1: *const u8, n1: usize, r2: *const u8,
                                     n2: usize, unit_mask_out: *mut u32, unit_mask_words: usize, out: *mut u8, out_cap: usize, out_len: *mut usize,
                                     out2: *mut u8, out2_cap: usize, out2_len: *mut usize, res: *mut FqgResult) -> c_int;
    pub fn fqg_grep_fastq_device(ctx: *mut c_void, m: *const c_void, mode: c_int, d_r1: *const u8, n1: usize, d_r2: *const u8, n2: usize,
                                 d_unit_mask_out: *mut u32, unit_mask_words: usize, res: *mut FqgResult, stream: *mut c_void) -> c_int;
    pub fn fqg_stream_open(ctx: *mut c_void, m: *const c_void, mode: c_int, opts: *const FqgStreamOpts, out: *mut *mut c_void) -> c_int;
    pub fn fqg_stream_acquire(s: *mut c_void, r1_buf: *mut *mut u8, r2_buf: *mut *mut u8, capacity: *mut usize) -> c_int;
    pub fn fqg_stream_submit(s: *mut c_void, n1: usize, n2: usize) -> c_int;
    pub fn fqg_stream_feed(s: *mut c_void, r1: *const u8, n1: usize, r2: *const u8, n2: usize) -> c_int;
    pub fn fqg_stream_in_flight(s: *const c_void) -> c_int;
    pub fn fqg_stream_next(s: *mut c_void, out: *mut FqgBatch) -> c_int;
    pub fn fqg_stream_close(s: *mut c_void);
    pub fn fqg_fastq_cut(mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize, limit: usize, is_final: c_int, exact: c_int,
                         cut1: *mut usize, cut2: *mut usize) -> c_int;
    pub fn fqg_write_spans(spans: *const FqgSpan, n_spans: u64, r1: *const u8, n1: usize, r2: *const u8, n2: usize, out: *mut u8, out_cap: usize,
                           out_len: *mut usize, out2: *mut u8, out2_cap: usize, out2_len: *mut usize) -> c_int;
    pub fn fqg_index_fastq_device(ctx: *mut c_void, d_fastq: *const u8, n: usize, d_seq_start: *mut u64, d_seq_len: *mut u32, capacity: u64,
                                  n_records: *mut u64, stream: *mut c_void) -> c_int;
    pub fn fqg_compact_bases_device(ctx: *mut c_void, d_fastq: *const u8, d_seq_start: *const u64, d_seq_len: *const u32, n_records: u64,
                                    d_bases_out: *mut u8, d_offsets_out: *mut u32, stream: *mut c_void) -> c_int;
    pub fn fqg_write_selected(mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize, unit_mask: *const u32,
                              out: *mut u8, out_len: *mut usize, out2: *mut u8, out2_len: *mut usize) -> c_int;
    pub fn fqg_read_input(path: *const c_char, decompress: c_int, pinned: c_int, out: *mut *mut u8, n_out: *mut usize) -> c_int;
    pub fn fqg_free_input(p: *mut u8, pinned: c_int);
    pub fn fqg_reader_open(path: *const c_char, decompress: c_int, out: *mut *mut c_void) -> c_int;
    pub fn fqg_reader_read(r: *mut c_void, dst: *mut u8, cap: usize, n_read: *mut usize, eof: *mut c_int) -> c_int;
    pub fn fqg_reader_close(r: *mut c_void);
    pub fn fqg_is_gzip_path(path: *const c_char) -> c_int;
    pub fn fqg_is_fastq_path(path: *const c_char) -> c_int;
    pub fn fqg_fastq_split(buf: *const u8, n: usize, parts: u32, cuts: *mut u64) -> c_int;
    pub fn fqg_synth_reads_device(ctx: *mut c_void, d_out: *mut u8, n_reads: u64, first_index: u64, read_len: u32, fastq: c_int, seed: u64,
                                  plant_blob: *const u8, plant_offsets: *const u32, n_plant: u32, plant_rate: f64, plant_seed: u64,
                                  stream: *mut c_void) -> c_int;
    pub fn fqg_kernel_launches() -> u64;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(fqg_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: c_int) -> Result<()> {
    if rc != 0 { bail!(last_error()); }
    Ok(())
}

/// validate_fixed_pattern (matcher.rs:128-147), same message text.
pub fn validate_fixed_pattern(pattern: &str, protein: bool) -> Result<()> {
    let c = CString::new(pattern)?;
    check(unsafe { fqg_validate_fixed_pattern(c.as_ptr(), protein as c_int) })
}

/// One device context per (process, GPU).
pub struct GpuContext(*mut c_void);
unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { fqg_ctx_create(device, &mut p) })?;
        Ok(Self(p))
    }
    pub fn device_count() -> i32 { unsafe { fqg_device_count() } }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { fqg_ctx_destroy(self.0) } } }

/// Base trait for all pattern matchers: the reference's trait (matcher.rs:150-193) without the colouring methods, which
/// stay on the host side of the reference (they need match positions, the device returns selection bits).
pub trait Matcher {
    fn opts(&self) -> MatcherOpts;
    /// true if the bases contain the pattern, XOR invert_match (matcher.rs:204, 250, 507, 570); no reverse complement
    fn bases_match(&self, bases: &[u8]) -> bool;
    /// (f(seq) || (reverse_complement && f(rc(seq)))) != invert_match (matcher.rs:164-175; :631-638 for read names)
    fn read_match(&self, read: &RefRecord) -> bool;
}

/// The batch seam (src/lib/seq_io.rs:314-326): one call per RecordSet instead of one virtual call per record.
pub trait BatchMatcher {
    /// bit k of word k / 32 = read_match(record k); `fastq` holds whole records (a RecordSet's buffer).
    fn match_batch(&self, fastq: &[u8], mate: Option<&[u8]>, mode: Mode) -> Result<(Vec<u32>, FqgResult)>;
}

/// The GPU stand-in for `Box<dyn Matcher + Sync + Send>`.
pub struct GpuMatcher { handle: *mut c_void, fwd_only: *mut c_void, kind: Kind, opts: MatcherOpts, ctx: std::sync::Arc<GpuContext> }
unsafe impl Send for GpuMatcher {}
unsafe impl Sync for GpuMatcher {}

fn new_handle<S: AsRef<[u8]>>(kind: Kind, patterns: &[S], opts: MatcherOpts) -> Result<*mut c_void> {
    let mut blob = Vec::new();
    let mut offs = vec![0u64];
    for p in patterns { blob.extend_from_slice(p.as_ref()); offs.push(blob.len() as u64); }
    let bits = (opts.invert_match as u32) | ((opts.reverse_complement as u32) << 1);
    let mut h = std::ptr::null_mut();
    check(unsafe { fqg_matcher_new(kind as c_int, blob.as_ptr(), offs.as_ptr(), patterns.len() as u32, bits, &mut h) })?;
    Ok(h)
}

impl GpuMatcher {
    pub fn new<S: AsRef<[u8]>>(ctx: std::sync::Arc<GpuContext>, kind: Kind, patterns: &[S], opts: MatcherOpts) -> Result<Self> {
        let handle = new_handle(kind, patterns, opts)?;
        // bases_match of the trait ignores reverse_complement: a second compiled set without it (cheap: host only)
        let fwd_only = if opts.reverse_complement && kind != Kind::Names {
            new_handle(kind, patterns, MatcherOpts { invert_match: opts.invert_match, reverse_complement: false })?
        } else { std::ptr::null_mut() };
        Ok(Self { handle, fwd_only, kind, opts, ctx })
    }
    pub fn handle(&self) -> *const c_void { self.handle }
    pub fn context(&self) -> &GpuContext { &self.ctx }

    /// read_match for every read of an SoA batch (bases back to back + offsets).
    pub fn match_bases(&self, bases: &[u8], offsets: &[u64]) -> Result<(Vec<u32>, u64)> {
        self.match_bases_with(self.handle, bases, offsets)
    }
    fn match_bases_with(&self, h: *mut c_void, bases: &[u8], offsets: &[u64]) -> Result<(Vec<u32>, u64)> {
        let n = offsets.len().saturating_sub(1) as u64;
        let mut mask = vec![0u32; (n as usize + 31) / 32 + 1];
        let mut sel = 0u64;
        check(unsafe { fqg_match_bases_host(self.ctx.0, h, bases.as_ptr(), offsets.as_ptr(), n, mask.as_mut_ptr(), &mut sel) })?;
        Ok((mask, sel))
    }
    /// Parse + match + ordered write-back of one host-resident input (src/main.rs:550-665 incl. 641-657, 682-727).
    pub fn grep_fastq_write(&self, r1: &[u8], r2: Option<&[u8]>, mode: Mode, out: &mut Vec<u8>) -> Result<FqgResult> {
        let (p2, n2) = r2.map(|b| (b.as_ptr(), b.len())).unwrap_or((std::ptr::null(), 0));
        out.resize(r1.len() + n2 + 64, 0);
        let mut n = 0usize;
        let mut res = FqgResult::default();
        check(unsafe { fqg_grep_fastq_host_write(self.ctx.0, self.handle, mode as c_int, r1.as_ptr(), r1.len(), p2, n2, out.as_mut_ptr(), out.len(), &mut n,
                                                 std::ptr::null_mut(), 0, std::ptr::null_mut(), &mut res) })?;
        out.truncate(n);
        Ok(res)
    }
}
impl Drop for GpuMatcher {
    fn drop(&mut self) { unsafe { fqg_matcher_free(self.handle); if !self.fwd_only.is_null() { fqg_matcher_free(self.fwd_only) } } }
}

impl Matcher for GpuMatcher {
    fn opts(&self) -> MatcherOpts { self.opts }
    fn bases_match(&self, bases: &[u8]) -> bool {
        if self.kind == Kind::Names { return !self.opts.invert_match; }            // matcher.rs:616-618
        let h = if self.fwd_only.is_null() { self.handle } else { self.fwd_only };
        let offs = [0u64, bases.len() as u64];
        self.match_bases_with(h, bases, &offs).map(|(_, n)| n == 1).expect("GPU match failed")
    }
    /// One record = a 1-record batch on the device: correct, and only meant for tests and `--color`'s second look at an
    /// already selected record (src/main.rs:707); the hot loop goes through BatchMatcher / FastqStream.
    fn read_match(&self, read: &RefRecord) -> bool {
        if self.kind == Kind::Names {
            let mut rec = Vec::with_capacity(read.head().len() + 2 * read.seq().len() + 8);
            rec.push(b'@'); rec.extend_from_slice(read.head()); rec.push(b'\n');
            rec.extend_from_slice(read.seq()); rec.extend_from_slice(b"\n+\n");
            rec.extend_from_slice(read.qual()); rec.push(b'\n');
            return self.match_batch(&rec, None, Mode::Single).map(|(_, r)| r.n_selected == 1).expect("GPU match failed");
        }
        let offs = [0u64, read.seq().len() as u64];
        self.match_bases(read.seq(), &offs).map(|(_, n)| n == 1).expect("GPU match failed")
    }
}

impl BatchMatcher for GpuMatcher {
    fn match_batch(&self, fastq: &[u8], mate: Option<&[u8]>, mode: Mode) -> Result<(Vec<u32>, FqgResult)> {
        let words = fastq.len() / 8 / 32 + 2;
        let mut mask = vec![0u32; words];
        let mut res = FqgResult::default();
        let (p2, n2) = mate.map(|b| (b.as_ptr(), b.len())).unwrap_or((std::ptr::null(), 0));
        check(unsafe { fqg_grep_fastq_host(self.ctx.0, self.handle, mode as c_int, fastq.as_ptr(), fastq.len(), p2, n2, mask.as_mut_ptr(), words, &mut res) })?;
        Ok((mask, res))
    }
}

/// The streaming batcher: a ring of batches in flight on one GPU, results in submission (= input) order.
pub struct FastqStream<'m> { h: *mut c_void, _m: &'m GpuMatcher }
impl<'m> FastqStream<'m> {
    pub fn open(m: &'m GpuMatcher, mode: Mode, opts: FqgStreamOpts) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        check(unsafe { fqg_stream_open(m.ctx.0, m.handle, mode as c_int, &opts, &mut h) })?;
        Ok(Self { h, _m: m })
    }
    /// The pinned buffers of the next free slot, for a reader thread to fill.
    pub fn acquire(&mut self) -> Result<(&mut [u8], Option<&mut [u8]>)> {
        let (mut p1, mut p2, mut cap) = (std::ptr::null_mut(), std::ptr::null_mut(), 0usize);
        check(unsafe { fqg_stream_acquire(self.h, &mut p1, &mut p2, &mut cap) })?;
        unsafe {
            let a = std::slice::from_raw_parts_mut(p1, cap);
            let b = if p2.is_null() { None } else { Some(std::slice::from_raw_parts_mut(p2, cap)) };
            Ok((a, b))
        }
    }
    pub fn submit(&mut self, n1: usize, n2: usize) -> Result<()> { check(unsafe { fqg_stream_submit(self.h, n1, n2) }) }
    pub fn in_flight(&self) -> usize { unsafe { fqg_stream_in_flight(self.h) as usize } }
    /// Result of the oldest batch in flight; its pointers stay valid until the slot is reused.
    pub fn next(&mut self) -> Result<FqgBatch> {
        let mut b = std::mem::MaybeUninit::<FqgBatch>::zeroed();
        check(unsafe { fqg_stream_next(self.h, b.as_mut_ptr()) })?;
        Ok(unsafe { b.assume_init() })
    }
}
impl Drop for FastqStream<'_> { fn drop(&mut self) { unsafe { fqg_stream_close(self.h) } } }

/// Factory for building a matcher (matcher.rs:641-686): same five arguments in the same order.  `bitencs` -- the
/// reference passes `Option<Vec<BitEnc>>` -- is the list of IUPAC pattern strings the BitEncs were made from
/// (`encode(pattern)`, src/main.rs:362-368): the 4-bit encoding happens inside the library.
pub struct MatcherFactory;
impl MatcherFactory {
    pub fn new_matcher(
        ctx: std::sync::Arc<GpuContext>,
        pattern: &Option<String>,
        fixed_strings: bool,
        regexp: &Vec<String>,
        bitencs: Option<Vec<String>>,
        match_opts: MatcherOpts,
    ) -> Result<Box<dyn Matcher + Sync + Send>> {
        let use_bitmask = bitencs.is_some();
        let is_single = pattern.is_some();
        let m = match (use_bitmask, is_single, fixed_strings) {
            (true, true, _) => GpuMatcher::new(ctx, Kind::BitMask, &bitencs.unwrap()[..1], match_opts)?,
            (true, false, _) => GpuMatcher::new(ctx, Kind::BitMaskSet, &bitencs.unwrap(), match_opts)?,
            (false, true, true) => GpuMatcher::new(ctx, Kind::Fixed, &[pattern.as_ref().unwrap()], match_opts)?,
            (false, true, false) => GpuMatcher::new(ctx, Kind::Regex, &[pattern.as_ref().unwrap()], match_opts)?,
            (false, false, true) => GpuMatcher::new(ctx, Kind::FixedSet, regexp, match_opts)?,
            (false, false, false) => GpuMatcher::new(ctx, Kind::RegexSet, regexp, match_opts)?,
        };
        Ok(Box::new(m))
    }

    /// Create a matcher for query names (read IDs), matcher.rs:679-685.
    pub fn new_query_name_matcher(ctx: std::sync::Arc<GpuContext>, query_names: AHashSet<Vec<u8>>, match_opts: MatcherOpts) -> Result<Box<dyn Matcher + Sync + Send>> {
        let v: Vec<Vec<u8>> = query_names.into_iter().collect();
        Ok(Box::new(GpuMatcher::new(ctx, Kind::Names, &v, match_opts)?))
    }
}
