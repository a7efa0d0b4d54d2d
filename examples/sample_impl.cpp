// sample_impl.cpp -- the reference's examples/sample_impl.rs (lines 1-128) written against the
// C++ mirror of its API (include/rcb200.hpp).  Same data, same table, same prints; every coding
// step runs on the GPU.  Exits non-zero if the round trip or the known answer fails.
#include <cstdio>
#include <vector>

#include "../include/rcb200.hpp"

using namespace range_coder;

// examples/sample_impl.rs:4-70
struct AlphabetParam {
    uint32_t cum;  // 文字の累積出現頻度 (cumulative frequency)
    uint32_t c;    // 文字の出現頻度 (frequency)
};
struct FreqTable : PModel {
    uint32_t total_freq_ = 0;
    std::vector<AlphabetParam> alphabet_params;
    explicit FreqTable(size_t alphabet_count) : alphabet_params(alphabet_count, AlphabetParam{0, 0}) {}
    uint32_t c_freq(size_t index) const override { return alphabet_params.at(index).c; }
    uint32_t cum_freq(size_t index) const override { return alphabet_params.at(index).cum; }
    uint32_t total_freq() const override { return total_freq_; }
    size_t alphabet_count() const override { return alphabet_params.size(); }
    void add_alphabet_freq(size_t alphabet_index) { alphabet_params[alphabet_index].c += 1; }
    void calc_cum() {
        uint32_t cum_total = 0;
        for (auto& a : alphabet_params) {
            a.cum = cum_total;
            cum_total += a.c;
        }
        total_freq_ = cum_total;
    }
};

int main() {
    // define test data (examples/sample_impl.rs:74)
    std::vector<size_t> test_data = {2, 1, 1, 4, 1, 4, 2, 1, 0, 1, 5, 9, 8, 7, 6, 5};

    // create freq-table
    FreqTable sd(10);
    for (size_t i : test_data) sd.add_alphabet_freq(i);
    sd.calc_cum();
    printf("FREQ TABLE\n");
    for (size_t i = 0; i < sd.alphabet_params.size(); i++)
        printf("index:%zu, c:%u, cum:%u\n", i, sd.c_freq(i), sd.cum_freq(i));
    printf("\n");

    // encode
    printf("ENCODING\nencode : ");
    Encoder encoder;
    uint32_t emitted = 0;
    for (size_t i : test_data) {
        printf("%zu,", i);
        emitted += encoder.encode(sd, i);
    }
    auto code = encoder.finish();
    printf("\noutput : 0x");
    for (uint8_t b : code) printf("%x", b);
    printf("\nlength : %zubyte\n\n", code.size());

    // decode
    Decoder decoder(code);
    printf("DECODING\ndecode : ");
    std::vector<size_t> decodeds;
    for (size_t k = 0; k < test_data.size(); k++) {
        size_t d = decoder.decode(sd);
        printf("%zu,", d);
        decodeds.push_back(d);
    }
    printf("\n\n");

    // test (examples/sample_impl.rs:123) + the known answer derived in SURVEY.md App. B.1
    const uint8_t expect[13] = {0x64, 0x47, 0x5f, 0x89, 0x70, 0x36, 0x5a, 0x2f, 0x83, 0xb2, 0x02, 0x46, 0xc0};
    bool ok = decodeds == test_data && code.size() == 13 && emitted == 5;
    for (size_t i = 0; ok && i < 13; i++) ok = code[i] == expect[i];

    // the same data through the bulk path: one chunk == one Encoder run
    Context& ctx = Context::thread_default();
    ModelSnapshot snap(ctx, sd);
    std::vector<uint8_t> syms(test_data.begin(), test_data.end());
    auto enc = gpu::encode_chunks(ctx, snap, syms, 16);
    ok = ok && enc.stream.size() == 13 && std::equal(enc.stream.begin(), enc.stream.end(), expect);
    auto back = gpu::decode_chunks<uint8_t>(ctx, snap, enc, syms.size(), 16);
    ok = ok && back == syms;

    printf(ok ? "test passed\n" : "TEST FAILED\n");
    return ok ? 0 : 1;
}
