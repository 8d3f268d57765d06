//! next-plaid/src/b200.rs -- binding of libplaid_b200 (include/plaid_b200.h) for the `b200` cargo
//! feature.  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no cargo/rustc.  It is the
//! shim a next-plaid maintainer adds so that `MmapIndex::{load, search, search_batch}` keep their
//! signatures (index.rs:1026, :1258, :1279) while the work runs on a B200; `colgrep` and
//! `next-plaid-api` link unchanged because they only see `MmapIndex`.
#![cfg(feature = "b200")]

use std::ffi::{c_char, c_void, CStr, CString};
use std::os::raw::c_int;

use ndarray::Array2;

use crate::error::{Error, Result};
use crate::search::{QueryResult, SearchParameters};

#[repr(C)]
pub struct PbSearchParams {
    pub batch_size: i64,
    pub n_full_scores: i64,
    pub top_k: i64,
    pub n_ivf_probe: i64,
    pub centroid_batch_size: i64,
    pub has_centroid_score_threshold: i32,
    pub centroid_score_threshold: f32,
}

impl From<&SearchParameters> for PbSearchParams {
    fn from(p: &SearchParameters) -> Self {
        PbSearchParams {
            batch_size: p.batch_size as i64,
            n_full_scores: p.n_full_scores as i64,
            top_k: p.top_k as i64,
            n_ivf_probe: p.n_ivf_probe as i64,
            centroid_batch_size: p.centroid_batch_size as i64,
            has_centroid_score_threshold: p.centroid_score_threshold.is_some() as i32,
            centroid_score_threshold: p.centroid_score_threshold.unwrap_or(0.0),
        }
    }
}

#[link(name = "plaid_b200")]
extern "C" {
    fn pb_index_load(index_dir: *const c_char, device: i32, out: *mut *mut c_void) -> c_int;
    fn pb_index_close(ix: *mut c_void);
    fn pb_index_embedding_dim(ix: *const c_void) -> i32;
    fn pb_search_batch(
        ix: *mut c_void,
        queries: *const f32,
        q_tok_offsets: *const i64,
        n_queries: i64,
        params: *const PbSearchParams,
        subset: *const i64,
        n_subset: i64,
        out_ids: *mut i64,
        out_scores: *mut f32,
        out_counts: *mut i32,
    ) -> c_int;
    fn pb_last_error() -> *const c_char;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(pb_last_error()).to_string_lossy().into_owned() }
}

/// Device-resident twin of the mmap'd arrays of `MmapIndex` (index.rs:995-1016).
pub struct B200Index {
    handle: *mut c_void,
}

// The C library is re-entrant on one handle (per-call stream + workspace pool), which is what
// `ArcSwap<MmapIndex>` + tokio workers need (next-plaid-api/src/state.rs:24-47).
unsafe impl Send for B200Index {}
unsafe impl Sync for B200Index {}

impl B200Index {
    /// Called at the end of `MmapIndex::load` (index.rs:1127): same directory, no conversion.
    pub fn load(index_path: &str, device: i32) -> Result<Self> {
        let c = CString::new(index_path).map_err(|e| Error::IndexLoad(e.to_string()))?;
        let mut handle: *mut c_void = std::ptr::null_mut();
        let st = unsafe { pb_index_load(c.as_ptr(), device, &mut handle) };
        if st != 0 {
            // No CPU fallback on this path: behaves like NEXT_PLAID_FORCE_GPU (lib.rs:71-84).
            return Err(Error::IndexLoad(last_error()));
        }
        Ok(B200Index { handle })
    }

    /// Body of `MmapIndex::search_batch` (index.rs:1279) under the `b200` feature.
    pub fn search_batch(
        &self,
        queries: &[Array2<f32>],
        params: &SearchParameters,
        subset: Option<&[i64]>,
    ) -> Result<Vec<QueryResult>> {
        // pb_search_batch has no dim argument and reads rows * embedding_dim floats: reject a query of another
        // width here, as the Python binding does (Error::Shape, not an out-of-bounds host read)
        let dim = unsafe { pb_index_embedding_dim(self.handle) } as usize;
        let mut offsets = Vec::with_capacity(queries.len() + 1);
        offsets.push(0i64);
        let mut flat: Vec<f32> = Vec::new();
        for q in queries {
            if q.ncols() != dim {
                return Err(Error::Shape(format!("query has {} columns, the index embedding_dim is {}", q.ncols(), dim)));
            }
            flat.extend(q.as_standard_layout().iter());
            offsets.push(offsets.last().unwrap() + q.nrows() as i64);
        }
        let k = params.top_k;
        let mut ids = vec![0i64; queries.len() * k.max(1)];
        let mut scores = vec![0f32; queries.len() * k.max(1)];
        let mut counts = vec![0i32; queries.len()];
        let p = PbSearchParams::from(params);
        let (sp, sn) = match subset {
            Some(s) if !s.is_empty() => (s.as_ptr(), s.len() as i64),
            Some(_) => (std::ptr::NonNull::<i64>::dangling().as_ptr() as *const i64, 0),
            None => (std::ptr::null(), 0),
        };
        let st = unsafe {
            pb_search_batch(
                self.handle,
                flat.as_ptr(),
                offsets.as_ptr(),
                queries.len() as i64,
                &p,
                sp,
                sn,
                ids.as_mut_ptr(),
                scores.as_mut_ptr(),
                counts.as_mut_ptr(),
            )
        };
        if st != 0 {
            return Err(Error::Search(last_error()));
        }
        Ok((0..queries.len())
            .map(|i| {
                let n = counts[i] as usize;
                QueryResult {
                    query_id: i, // search.rs:661
                    passage_ids: ids[i * k..i * k + n].to_vec(),
                    scores: scores[i * k..i * k + n].to_vec(),
                }
            })
            .collect())
    }

    /// Body of `MmapIndex::search` (index.rs:1258).
    pub fn search(
        &self,
        query: &Array2<f32>,
        params: &SearchParameters,
        subset: Option<&[i64]>,
    ) -> Result<QueryResult> {
        let mut r = self.search_batch(std::slice::from_ref(query), params, subset)?;
        let mut r = r.remove(0);
        r.query_id = 0;
        Ok(r)
    }
}

impl Drop for B200Index {
    fn drop(&mut self) {
        unsafe { pb_index_close(self.handle) }
    }
}
