// multi_gpu.cpp -- a static frequency table shared by chunks that live on several GPUs, driven by
// one host thread through the C++ mirror (include/rcb200.hpp: gpu::MultiGpu).  The caller-side loop it
// replaces is the reference's examples/sample_impl.rs:77-81 (count every symbol, calc_cum) followed
// by :92-98 and :113-120 (encode / decode loops), run once per 64 KiB chunk.
// Check: the sharded stream is byte-identical to the single-GPU stream of the same symbols under the
// same table, and decodes back.  Usage: multi_gpu [n_gpus] (default: all visible, at most 8).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/rcb200.hpp"

using namespace range_coder;

namespace {
// FreqTable of the reference's example, filled from raw counts
struct CountTable : PModel {
    std::vector<uint32_t> c, cum;
    uint32_t total = 0;
    explicit CountTable(const std::vector<uint32_t>& counts) : c(counts), cum(counts.size()) {
        for (size_t i = 0; i < c.size(); i++) {  // calc_cum, examples/sample_impl.rs:61-69
            cum[i] = total;
            total += c[i];
        }
    }
    uint32_t c_freq(size_t i) const override { return c.at(i); }
    uint32_t cum_freq(size_t i) const override { return cum.at(i); }
    uint32_t total_freq() const override { return total; }
    size_t alphabet_count() const override { return c.size(); }
};
}  // namespace

int main(int argc, char** argv) {
    int n_gpus = argc > 1 ? atoi(argv[1]) : rcb_device_count();
    if (n_gpus > 8) n_gpus = 8;
    if (n_gpus < 1) {
        fprintf(stderr, "no CUDA device (there is no CPU fallback)\n");
        return 2;
    }
    const uint32_t K = 256;
    const uint64_t chunk = 65536, n = 37 * chunk + 12345;  // ragged last chunk, uneven shards
    std::vector<uint8_t> syms(n);
    uint64_t x = 0x5EED0001;
    std::vector<uint32_t> counts(K, 0);
    for (uint64_t i = 0; i < n; i++) {  // skewed bytes: min of two uniform draws
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const uint32_t a = (uint32_t)(x >> 33) & 255u, b = (uint32_t)(x >> 41) & 255u;
        syms[i] = (uint8_t)(a < b ? a : b);
        counts[syms[i]]++;
    }
    gpu::MultiGpu multi(n_gpus);
    auto sharded = multi.encode_chunks(syms, K, chunk);
    auto back = multi.decode_chunks<uint8_t>(sharded, n, chunk);

    Context& ctx = Context::thread_default();
    CountTable table(counts);  // the table of the whole data, built the reference's way
    ModelSnapshot snap(ctx, table);
    auto single = gpu::encode_chunks(ctx, snap, syms, chunk);

    bool ok = back == syms && sharded.stream == single.stream && sharded.offsets == single.offsets;
    // every GPU must hold the identical table after the all-reduce
    for (int g = 0; ok && g < n_gpus; g++) {
        std::vector<uint32_t> c(K), cum(K);
        uint32_t total = 0;
        check(rcb_model_get_tables(multi.context(g), multi.model(g), 0, c.data(), cum.data(), &total, nullptr),
              "rcb_model_get_tables");
        ok = total == table.total && c == table.c && cum == table.cum;
    }
    printf("%d GPU(s): %llu symbols in %zu chunks -> %zu bytes; sharded stream %s the single-GPU stream\n", n_gpus,
           (unsigned long long)n, sharded.offsets.size() - 1, sharded.stream.size(), ok ? "==" : "!=");
    printf(ok ? "test passed\n" : "TEST FAILED\n");
    return ok ? 0 : 1;
}
