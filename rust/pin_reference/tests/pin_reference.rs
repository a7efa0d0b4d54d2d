//! Known-answer test of `tests/golden/vectors.json` against the unmodified `range_coder` crate.
//!
//! The vectors were produced by this repository's Python transliteration of the crate
//! (`tests/golden/make_golden.py`); the CPU oracle (`oracle/rc_oracle.c`) and the CUDA kernels are
//! tested against the same file.  This test closes the loop: the reference's own `Encoder::encode` /
//! `Encoder::finish` must produce exactly those bytes, and its `Decoder::decode` must return the
//! symbols; where a vector carries restart records (`restart`), the Encoder's `lower_bound`, `range` and
//! code length in front of those symbols must be the recorded ones.  The frequency table below is a plain `PModel` over the vector's `c` / `cum` / `total`
//! with the lookup rule of the reference's example model (examples/sample_impl.rs:27-45).
use range_coder::{Decoder, Encoder, PModel};
use serde_json::Value;
use sha2::{Digest, Sha256};
use std::collections::VecDeque;

struct Table {
    c: Vec<u32>,
    cum: Vec<u32>,
    total: u32,
}

impl PModel for Table {
    fn c_freq(&self, index: usize) -> u32 {
        self.c[index]
    }
    fn cum_freq(&self, index: usize) -> u32 {
        self.cum[index]
    }
    fn total_freq(&self) -> u32 {
        self.total
    }
    fn find_index(&self, decoder: &Decoder) -> usize {
        let rfreq = (decoder.data() - decoder.range_coder().lower_bound())
            / decoder.range_coder().range_par_total(self.total_freq());
        let (mut left, mut right) = (0usize, self.c.len() - 1);
        while left < right {
            let mid = (left + right) / 2;
            if self.cum_freq(mid + 1) as u64 <= rfreq {
                left = mid + 1;
            } else {
                right = mid;
            }
        }
        left
    }
}

fn u32s(v: &Value) -> Vec<u32> {
    v.as_array().unwrap().iter().map(|x| x.as_u64().unwrap() as u32).collect()
}

fn symbols_of(v: &Value, k: usize) -> Vec<usize> {
    if let Some(a) = v.get("symbols") {
        return a.as_array().unwrap().iter().map(|x| x.as_u64().unwrap() as usize).collect();
    }
    if let Some(r) = v.get("symbols_repeat") {
        let r = r.as_array().unwrap();
        return vec![r[0].as_u64().unwrap() as usize; r[1].as_u64().unwrap() as usize];
    }
    let raw = hex::decode(v["symbols_hex"].as_str().unwrap()).unwrap();
    if k > 256 {
        raw.chunks(2).map(|p| p[0] as usize | (p[1] as usize) << 8).collect() // little-endian u16
    } else {
        raw.iter().map(|&b| b as usize).collect()
    }
}

fn table_of(v: &Value, k: usize, symbols: &[usize]) -> Table {
    if v.get("c_from_symbols").and_then(|b| b.as_bool()).unwrap_or(false) {
        // FreqTable::add_alphabet_freq + calc_cum (examples/sample_impl.rs:58-69)
        let mut c = vec![0u32; k];
        for &s in symbols {
            c[s] += 1;
        }
        let mut cum = vec![0u32; k];
        let mut run = 0u32;
        for i in 0..k {
            cum[i] = run;
            run += c[i];
        }
        return Table { c, cum, total: run };
    }
    Table { c: u32s(&v["c"]), cum: u32s(&v["cum"]), total: v["total"].as_u64().unwrap() as u32 }
}

#[test]
fn golden_vectors_match_the_reference_crate() {
    let path = concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/vectors.json");
    let vectors: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let mut written = Vec::new();
    for v in vectors.as_array().unwrap() {
        let name = v["name"].as_str().unwrap();
        let k = v["K"].as_u64().unwrap() as usize;
        let symbols = symbols_of(v, k);
        assert_eq!(symbols.len() as u64, v["n_symbols"].as_u64().unwrap(), "{}", name);
        let table = table_of(v, k, &symbols);
        assert_eq!(table.total as u64, v["total"].as_u64().unwrap(), "{}", name);

        // the caller's loop of examples/sample_impl.rs:92-98
        let mut encoder = Encoder::new();
        let mut produced = 0u64;
        // restart points (include/rcb200.h: rcb_restart_point): the Encoder's own state in front of every
        // restart_syms-th symbol -- RangeCoder::lower_bound() / range() (src/range_coder.rs:28-35) through the
        // public field Encoder::range_coder (src/encoder.rs:8), and peek_code().len() (src/encoder.rs:15-17)
        let restart = v.get("restart");
        let rs = restart.map(|r| r["restart_syms"].as_u64().unwrap() as usize).unwrap_or(0);
        let mut seen = 0usize;
        for (j, &s) in symbols.iter().enumerate() {
            if rs != 0 && j != 0 && j % rs == 0 {
                let rec = restart.unwrap()["records"][seen].as_array().unwrap();
                let lower = u64::from_str_radix(rec[0].as_str().unwrap(), 16).unwrap();
                let range = u64::from_str_radix(rec[1].as_str().unwrap(), 16).unwrap();
                assert_eq!(encoder.range_coder.lower_bound(), lower, "{}: restart record {} lower_bound", name, seen);
                assert_eq!(encoder.range_coder.range(), range, "{}: restart record {} range", name, seen);
                assert_eq!(encoder.peek_code().len() as u64, rec[2].as_u64().unwrap(), "{}: restart record {} bytes", name, seen);
                assert_eq!(produced, rec[2].as_u64().unwrap(), "{}: encode() return values add up", name);
                seen += 1;
            }
            produced += encoder.encode(&table, s) as u64;
        }
        if let Some(r) = restart {
            assert_eq!(seen, r["records"].as_array().unwrap().len(), "{}: every restart record checked", name);
        }
        let code: VecDeque<u8> = encoder.finish();
        let code: Vec<u8> = code.into_iter().collect();
        assert_eq!(code.len() as u64, produced + 8, "{}: finish() appends 8 bytes", name);
        assert_eq!(code.len() as u64, v["code_len"].as_u64().unwrap(), "{}: code length", name);
        let sha = hex::encode(Sha256::digest(&code));
        assert_eq!(sha, v["code_sha256"].as_str().unwrap(), "{}: code bytes (sha256)", name);
        if let Some(h) = v.get("code_hex") {
            assert_eq!(hex::encode(&code), h.as_str().unwrap(), "{}: code bytes", name);
        }

        // and back: examples/sample_impl.rs:110-120
        let mut decoder = Decoder::new(code.clone());
        let decoded: Vec<usize> = symbols.iter().map(|_| decoder.decode(&table)).collect();
        assert_eq!(decoded, symbols, "{}: round trip", name);
        println!("pinned {:24} {:6} symbols -> {:5} bytes  sha256 {}", name, symbols.len(), code.len(), &sha[..16]);
        written.push(serde_json::json!({"name": name, "code_len": code.len(), "code_sha256": sha,
                                        "code_hex": hex::encode(&code), "generator": "range_coder crate (unmodified)"}));
    }
    if let Ok(out) = std::env::var("RCB_PIN_WRITE") {
        std::fs::write(out, serde_json::to_string_pretty(&Value::Array(written)).unwrap()).unwrap();
    }
}

/// a table the caller updates between calls (`&T` per call, src/encoder.rs:24) -- counts start
/// at 1, the coded symbol gains `inc` after it was coded, all counts are halved (rounding up) when the total
/// would pass `limit`; calc_cum after every change (examples/sample_impl.rs:61-69).
struct AdaptiveTable {
    t: Table,
    inc: u32,
    limit: u32,
}

impl AdaptiveTable {
    fn new(k: usize, inc: u32, limit: u32) -> Self {
        let mut a = AdaptiveTable { t: Table { c: vec![1u32; k], cum: vec![0u32; k], total: 0 }, inc, limit };
        a.calc_cum();
        a
    }
    fn calc_cum(&mut self) {
        let mut run = 0u32;
        for i in 0..self.t.c.len() {
            self.t.cum[i] = run;
            run += self.t.c[i];
        }
        self.t.total = run;
    }
    fn update(&mut self, s: usize) {
        self.t.c[s] += self.inc;
        if self.t.total as u64 + self.inc as u64 > self.limit as u64 {
            for x in self.t.c.iter_mut() {
                *x = (*x + 1) >> 1;
            }
        }
        self.calc_cum();
    }
}

#[test]
fn adaptive_vectors_match_the_reference_crate() {
    let path = concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/adaptive_vectors.json");
    let vectors: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    for v in vectors.as_array().unwrap() {
        let name = v["name"].as_str().unwrap();
        let k = v["K"].as_u64().unwrap() as usize;
        let inc = v["adaptive"]["inc"].as_u64().unwrap() as u32;
        let limit = v["adaptive"]["limit"].as_u64().unwrap() as u32;
        let symbols = symbols_of(v, k);
        let mut table = AdaptiveTable::new(k, inc, limit);
        let mut encoder = Encoder::new();
        for &s in &symbols {
            encoder.encode(&table.t, s); // the table as it is before the update
            table.update(s);
        }
        let code: Vec<u8> = encoder.finish().into_iter().collect();
        assert_eq!(code.len() as u64, v["code_len"].as_u64().unwrap(), "{}: code length", name);
        assert_eq!(hex::encode(Sha256::digest(&code)), v["code_sha256"].as_str().unwrap(), "{}: code bytes", name);
        let mut table = AdaptiveTable::new(k, inc, limit);
        let mut decoder = Decoder::new(code);
        for &s in &symbols {
            let got = decoder.decode(&table.t);
            assert_eq!(got, s, "{}: round trip", name);
            table.update(got);
        }
        println!("pinned {:24} {:6} symbols (adaptive table)", name, symbols.len());
    }
}
