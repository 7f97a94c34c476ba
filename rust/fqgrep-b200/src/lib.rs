//! Rust shim over the C ABI in include/fqgrep_b200.h.  Mirrors the reference's `MatcherOpts` / `MatcherFactory`
//! surface (fqgrep src/lib/matcher.rs:20-28, 641-686) and adds the batch-level call the GPU needs.
//! NOTE: shipped as source only; the build image has no cargo/rustc, so this file is compile-unverified.
use anyhow::{bail, Result};
use std::ffi::{c_char, c_int, c_void, CStr, CString};

#[derive(Copy, Clone, Debug)]
pub struct MatcherOpts {
    pub invert_match: bool,
    pub reverse_complement: bool,
}

#[repr(i32)]
#[derive(Copy, Clone, Debug)]
pub enum Kind { Fixed = 0, FixedSet = 1, Regex = 2, RegexSet = 3, Names = 4, BitMask = 5, BitMaskSet = 6 }

#[repr(i32)]
#[derive(Copy, Clone, Debug)]
pub enum Mode { Single = 0, Interleaved = 1, Paired = 2 }

#[repr(C)]
#[derive(Default, Debug)]
pub struct FqgResult {
    pub n_records: u64, pub n_units: u64, pub n_selected: u64,
    pub h2d_ms: f64, pub kernel_ms: f64, pub d2h_ms: f64,
    pub h2d_bytes: u64, pub d2h_bytes: u64,
}

#[link(name = "fqgrep_b200")]
extern "C" {
    fn fqg_last_error() -> *const c_char;
    fn fqg_matcher_new(kind: c_int, blob: *const u8, offsets: *const u64, n: u32, opts: u32, out: *mut *mut c_void) -> c_int;
    fn fqg_matcher_free(m: *mut c_void);
    fn fqg_validate_fixed_pattern(pattern: *const c_char, protein: c_int) -> c_int;
    fn fqg_ctx_create(device: c_int, out: *mut *mut c_void) -> c_int;
    fn fqg_ctx_destroy(ctx: *mut c_void);
    fn fqg_match_bases_host(ctx: *mut c_void, m: *const c_void, bases: *const u8, offsets: *const u64, n_reads: u64,
                            mask_out: *mut u32, n_selected: *mut u64) -> c_int;
    fn fqg_grep_fastq_host(ctx: *mut c_void, m: *const c_void, mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize,
                           unit_mask_out: *mut u32, unit_mask_words: usize, res: *mut FqgResult) -> c_int;
    fn fqg_write_selected(mode: c_int, r1: *const u8, n1: usize, r2: *const u8, n2: usize, unit_mask: *const u32,
                          out: *mut u8, out_len: *mut usize, out2: *mut u8, out2_len: *mut usize) -> c_int;
    // The rest of include/fqgrep_b200.h, declared so the whole C ABI is reachable from Rust (used by callers that keep
    // their reads resident on the device, read inputs through the library, or shard across GPUs).
    pub fn fqg_version() -> *const c_char;
    pub fn fqg_device_count() -> c_int;
    pub fn fqg_kernel_launches() -> u64;
    pub fn fqg_iupac_expand(pattern: *const c_char, to_regex: c_int, out: *mut c_char, cap: usize, out_len: *mut usize) -> c_int;
    pub fn fqg_matcher_dfa_info(m: *const c_void, out: *mut FqgDfaInfo) -> c_int;
    pub fn fqg_matcher_dfa_export(m: *const c_void, byte_class: *mut u8, next: *mut u32) -> c_int;
    pub fn fqg_alloc_pinned(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn fqg_free_pinned(p: *mut c_void);
    pub fn fqg_match_bases_device(ctx: *mut c_void, m: *const c_void, d_bases: *const u8, d_start: *const u32, d_end: *const u32, n_reads: u64,
                                  d_mask_out: *mut u32, d_or_mask: *const u32, d_count: *mut u64, avg_read_len: u32, stream: *mut c_void) -> c_int;
    pub fn fqg_grep_fastq_device(ctx: *mut c_void, m: *const c_void, mode: c_int, d_r1: *const u8, n1: usize, d_r2: *const u8, n2: usize,
                                 d_unit_mask_out: *mut u32, unit_mask_words: usize, res: *mut FqgResult, stream: *mut c_void) -> c_int;
    pub fn fqg_index_fastq_device(ctx: *mut c_void, d_fastq: *const u8, n: usize, d_seq_start: *mut u64, d_seq_len: *mut u32, capacity: u64,
                                  n_records: *mut u64, stream: *mut c_void) -> c_int;
    pub fn fqg_compact_bases_device(ctx: *mut c_void, d_fastq: *const u8, d_seq_start: *const u64, d_seq_len: *const u32, n_records: u64,
                                    d_bases_out: *mut u8, d_offsets_out: *mut u32, stream: *mut c_void) -> c_int;
    pub fn fqg_read_input(path: *const c_char, decompress: c_int, pinned: c_int, out: *mut *mut u8, n_out: *mut usize) -> c_int;
    pub fn fqg_free_input(p: *mut u8, pinned: c_int);
    pub fn fqg_is_gzip_path(path: *const c_char) -> c_int;
    pub fn fqg_is_fastq_path(path: *const c_char) -> c_int;
    pub fn fqg_fastq_split(buf: *const u8, n: usize, parts: u32, cuts: *mut u64) -> c_int;
    pub fn fqg_synth_reads_device(ctx: *mut c_void, d_out: *mut u8, n_reads: u64, first_index: u64, read_len: u32, fastq: c_int, seed: u64,
                                  plant_blob: *const u8, plant_offsets: *const u32, n_plant: u32, plant_rate: f64, plant_seed: u64,
                                  stream: *mut c_void) -> c_int;
}

/// fqg_dfa_info of the C header (compiled-automaton introspection).
#[repr(C)]
#[derive(Default, Debug)]
pub struct FqgDfaInfo {
    pub n_states: u32, pub n_classes: u32, pub start: u32, pub accept_from: u32, pub tier: u32, pub table_bytes: u32,
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(fqg_last_error()).to_string_lossy().into_owned() }
}

/// validate_fixed_pattern (matcher.rs:128-147), same message text.
pub fn validate_fixed_pattern(pattern: &str, protein: bool) -> Result<()> {
    let c = CString::new(pattern)?;
    if unsafe { fqg_validate_fixed_pattern(c.as_ptr(), protein as c_int) } != 0 { bail!(last_error()); }
    Ok(())
}

pub struct GpuContext(*mut c_void);
unsafe impl Send for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self> {
        let mut p = std::ptr::null_mut();
        if unsafe { fqg_ctx_create(device, &mut p) } != 0 { bail!(last_error()); }
        Ok(Self(p))
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { fqg_ctx_destroy(self.0) } } }

/// The GPU stand-in for `Box<dyn Matcher + Sync + Send>`.
pub struct GpuMatcher { handle: *mut c_void, opts: MatcherOpts }
unsafe impl Send for GpuMatcher {}
unsafe impl Sync for GpuMatcher {}

impl GpuMatcher {
    pub fn new<S: AsRef<[u8]>>(kind: Kind, patterns: &[S], opts: MatcherOpts) -> Result<Self> {
        let mut blob = Vec::new();
        let mut offs = vec![0u64];
        for p in patterns { blob.extend_from_slice(p.as_ref()); offs.push(blob.len() as u64); }
        let bits = (opts.invert_match as u32) | ((opts.reverse_complement as u32) << 1);
        let mut h = std::ptr::null_mut();
        let rc = unsafe { fqg_matcher_new(kind as c_int, blob.as_ptr(), offs.as_ptr(), patterns.len() as u32, bits, &mut h) };
        if rc != 0 { bail!(last_error()); }
        Ok(Self { handle: h, opts })
    }
    pub fn opts(&self) -> MatcherOpts { self.opts }

    /// read_match for every read of an SoA batch (replaces the loop at src/lib/seq_io.rs:314-326).
    pub fn match_bases(&self, ctx: &GpuContext, bases: &[u8], offsets: &[u64]) -> Result<(Vec<u32>, u64)> {
        let n = offsets.len().saturating_sub(1) as u64;
        let mut mask = vec![0u32; (n as usize + 31) / 32 + 1];
        let mut sel = 0u64;
        let rc = unsafe { fqg_match_bases_host(ctx.0, self.handle, bases.as_ptr(), offsets.as_ptr(), n, mask.as_mut_ptr(), &mut sel) };
        if rc != 0 { bail!(last_error()); }
        Ok((mask, sel))
    }

    /// Parse + match raw FASTQ (replaces the per-input body of src/main.rs:550-665).
    pub fn match_fastq(&self, ctx: &GpuContext, r1: &[u8], r2: Option<&[u8]>, mode: Mode) -> Result<(Vec<u32>, FqgResult)> {
        let words = r1.len() / 8 / 32 + 2;
        let mut mask = vec![0u32; words];
        let mut res = FqgResult::default();
        let (p2, n2) = r2.map(|b| (b.as_ptr(), b.len())).unwrap_or((std::ptr::null(), 0));
        let rc = unsafe { fqg_grep_fastq_host(ctx.0, self.handle, mode as c_int, r1.as_ptr(), r1.len(), p2, n2, mask.as_mut_ptr(), words, &mut res) };
        if rc != 0 { bail!(last_error()); }
        Ok((mask, res))
    }
}
impl Drop for GpuMatcher { fn drop(&mut self) { unsafe { fqg_matcher_free(self.handle) } } }

/// Ordered write-back of selected units (src/main.rs:641-657, 682-727).
pub fn write_selected(mode: Mode, r1: &[u8], r2: Option<&[u8]>, mask: &[u32]) -> Result<Vec<u8>> {
    let (p2, n2) = r2.map(|b| (b.as_ptr(), b.len())).unwrap_or((std::ptr::null(), 0));
    let mut out = vec![0u8; r1.len() + n2 + 64];
    let mut n = 0usize;
    let rc = unsafe { fqg_write_selected(mode as c_int, r1.as_ptr(), r1.len(), p2, n2, mask.as_ptr(), out.as_mut_ptr(), &mut n, std::ptr::null_mut(), std::ptr::null_mut()) };
    if rc != 0 { bail!(last_error()); }
    out.truncate(n);
    Ok(out)
}

/// MatcherFactory (matcher.rs:641-686); bit-mask variants are not part of the B200 path yet.
pub struct MatcherFactory;
impl MatcherFactory {
    pub fn new_matcher(pattern: &Option<String>, fixed_strings: bool, regexp: &Vec<String>, match_opts: MatcherOpts) -> Result<GpuMatcher> {
        match (pattern, fixed_strings) {
            (Some(p), true) => GpuMatcher::new(Kind::Fixed, &[p.as_bytes()], match_opts),
            (Some(p), false) => GpuMatcher::new(Kind::Regex, &[p.as_bytes()], match_opts),
            (None, true) => GpuMatcher::new(Kind::FixedSet, regexp, match_opts),
            (None, false) => GpuMatcher::new(Kind::RegexSet, regexp, match_opts),
        }
    }
    pub fn new_query_name_matcher<I: IntoIterator<Item = Vec<u8>>>(query_names: I, match_opts: MatcherOpts) -> Result<GpuMatcher> {
        let v: Vec<Vec<u8>> = query_names.into_iter().collect();
        GpuMatcher::new(Kind::Names, &v, match_opts)
    }
}
