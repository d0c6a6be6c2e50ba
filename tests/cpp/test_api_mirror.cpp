// The reference's own round-trip tests (`prove_verify_works` of suite_tests! / ietf_suite_tests! / pedersen_suite_tests!,
// SURVEY.md 4) replayed through the C++ mirror of its API (include/vrfs_b200.hpp), plus upstream Bandersnatch vector 1.
// Built and run by tests/test_gpu_cpp_mirror.py on a GPU box; `--compile-only` use on CPU boxes is just the build.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/vrfs_b200.hpp"

using namespace vrfs;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static Bytes bytes(const std::string& s) { return Bytes(s.begin(), s.end()); }
static Bytes unhex(const std::string& h) {
  Bytes o;
  for (size_t i = 0; i + 1 < h.size(); i += 2) o.push_back((uint8_t)std::strtoul(h.substr(i, 2).c_str(), nullptr, 16));
  return o;
}
static std::string hex(const uint8_t* p, size_t n) {
  static const char* d = "0123456789abcdef"; std::string s;
  for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; }
  return s;
}
static bool all(const Bytes& v, uint8_t x) { for (auto b : v) if (b != x) return false; return !v.empty(); }

static void prove_verify_works(const Suite& suite, const char* name) {
  const size_t n = 48;
  std::vector<Bytes> seeds, alphas, ad, ad2;
  for (size_t i = 0; i < n; i++) { seeds.push_back(bytes("TEST_SEED-" + std::to_string(i))); alphas.push_back(bytes("foo-" + std::to_string(i))); ad.push_back(bytes("bar")); ad2.push_back(bytes("baz")); }
  Secret secret = Secret::from_seed(suite, seeds);
  Public pub = secret.public_();
  auto [input, ok] = Input::new_(suite, alphas);
  CHECK(all(ok, 1));
  Output output = secret.output(input);
  ietf::Proof proof = secret.prove(input, output, ad);
  CHECK(all(pub.verify(input, output, ad, proof), 1));
  Errors errs;
  CHECK(all(pub.verify(input, output, ad2, proof, &errs), 0));                  // Error::VerificationFailure
  for (auto e : errs) CHECK(e == Error::VerificationFailure);
  {   // a value no typed Input can hold (off the curve) is Error::InvalidData, and only that item
    Input bad = input; bad.xy[5] ^= 0x10;
    Bytes okb = pub.verify(bad, output, ad, proof, &errs);
    CHECK(okb[0] == 0 && errs[0] == Error::InvalidData && okb[1] == 1 && errs[1] == Error::None);
  }
  auto [ped, blinding] = secret.pedersen_prove(input, output, ad);
  CHECK(all(pedersen::verify(suite, input, output, ad, ped), 1));
  CHECK(all(pedersen::verify(suite, input, output, ad2, ped, &errs), 0));
  for (auto e : errs) CHECK(e == Error::VerificationFailure);
  {   // the serialised proof (160 B for the 32-byte codecs): same blinding, verifies, and a flipped byte of `s` fails only its item
    auto [sp, bl2] = secret.pedersen_prove_serialized(input, output, ad);
    CHECK(bl2 == blinding && sp.bytes.size() == n * suite.pedersen_proof_len());
    CHECK(all(pedersen::verify(suite, input, output, ad, sp), 1));
    sp.bytes[3 * suite.point_enc_len() + 7] ^= 1;
    Bytes okp = pedersen::verify(suite, input, output, ad, sp, &errs);
    CHECK(okp[0] == 0 && okp[1] == 1 && errs[0] != Error::None && errs[1] == Error::None);
  }
  CHECK(blinding.size() == 32 * n && output.hash().size() == suite.hash_len() * n);
  // serialisation: keys round-trip with validation; signatures straight off the wire
  auto [pub2, kok] = Public::deserialize_compressed(suite, pub.serialize_compressed());
  CHECK(all(kok, 1) && pub2.xy == pub.xy);
  auto [sig, sok] = secret.sign(alphas, ad);
  CHECK(all(sok, 1) && sig.size() == n * suite.ietf_signature_len());
  CHECK(std::memcmp(sig.data(), output.serialize_compressed().data(), suite.point_enc_len()) == 0);
  auto [vok, beta] = Public::verify_signatures(suite, pub.serialize_compressed(), alphas, sig, ad);
  CHECK(all(vok, 1) && beta == output.hash());
  sig[suite.point_enc_len() + 3] ^= 1;                                           // corrupt item 0's challenge
  auto [vok2, beta2] = Public::verify_signatures(suite, pub.serialize_compressed(), alphas, sig, ad);
  CHECK(vok2[0] == 0 && vok2[1] == 1 && all(Bytes(beta2.begin(), beta2.begin() + suite.hash_len()), 0));
  std::printf("%s: prove_verify_works %s\n", name, failures ? "FAILED" : "ok");
}

int main() {
  try {
    Engine eng(0);
    prove_verify_works(Suite::bandersnatch(eng), "bandersnatch");
    prove_verify_works(Suite::ed25519(eng), "ed25519");
    prove_verify_works(Suite::secp256r1(eng), "secp256r1");
    // upstream bandersnatch_sha-512_ell2_ietf vector 1 (tests/golden/bandersnatch_upstream.json): seed = [01], alpha = "", ad = ""
    Suite s = Suite::bandersnatch(eng);
    Secret k = Secret::from_seed(s, {Bytes{0x01}});
    CHECK(hex(k.scalars.data(), 32) == "3d6406500d4009fdf2604546093665911e753f2213570a29521fd88bc30ede18");
    CHECK(hex(k.public_().serialize_compressed().data(), 32) == "a1b1da71cc4682e159b7da23050d8b6261eb11a3247c89b07ef56ccd002fd38b");
    auto [sig, ok] = k.sign({Bytes{}}, {Bytes{}});
    CHECK(ok[0] == 1);
    CHECK(hex(sig.data(), 96) == "e7aa5154103450f0a0525a36a441f827296ee489ef30ed8787cff8df1bef223f"
                                 "439fd9495643314fa623f2581f4b3d7d6037394468084f4ad7d8031479d9d101"
                                 "828bedd2ad95380b11f67a05ea0a76f0c3fef2bee9f043f4dffdddde09f55c01");
    HostRegistration pinned_sig(sig);                    // the signature bytes page-locked in place for the verify call
    CHECK(pinned_sig.ok());
    auto [vok, beta] = Public::verify_signatures(s, k.public_().serialize_compressed(), {Bytes{}}, sig, {Bytes{}});
    CHECK(vok[0] == 1);
    CHECK(hex(beta.data(), 64) == "fdeb377a4ffd7f95ebe48e5b43a88d069ce62188e49493500315ad55ee04d7442b93c4c91d5475370e9380496f4bc0b838c2483bce4e133c6f18b0adbb9e4722");
    // RingContext mirror: the one-call commitment equals the MSM of the columns it builds (SRS: multiples of one valid point -
    // any G1 points serve for this identity; here N copies of the BLS12-381 G1 generator), and the FFT round-trips
    {
      const size_t N = 64, NK = 20;
      Bytes g1 = unhex("bbc622db0af03afbef1a7af93fe8556c58ac1b173f3a4ea105b974974f8c68c30faca94f8c63952694d79731a7d3f117"
                       "e1e7c5462923aa0ce48a88a244c73cd0edb3042ccb18db00f60ad0d595e0f5fce48a1d74ed309ea0f1a0aae381f4b308");
      Bytes srs; for (size_t i = 0; i < N; i++) srs.insert(srs.end(), g1.begin(), g1.end());
      std::vector<Bytes> seeds; for (size_t i = 0; i < NK + 6; i++) seeds.push_back(bytes("ring-" + std::to_string(i)));
      Bytes pts = Secret::from_seed(s, seeds).public_points;
      Bytes keys(pts.begin(), pts.begin() + 64 * NK), padding(pts.begin() + 64 * NK, pts.begin() + 64 * (NK + 1)), tail(pts.begin() + 64 * (NK + 1), pts.end());
      RingContext rc(eng, srs, true, padding, tail);
      CHECK(rc.max_ring_size() == N - 3 - 5 - 1);
      Bytes cols = rc.fixed_columns(keys), com = rc.verifier_key_commitment(keys);
      CHECK(std::memcmp(cols.data(), keys.data(), 32) == 0 && cols[2 * N * 32] == 1 && cols[(3 * N - 1) * 32] == 0);
      CHECK(com == ring_commitment_msm(eng, srs, cols, 3));
      {   // homomorphic form: commitment of the padding-only ring + delta over the keys = the full commitment
        Bytes empty = rc.verifier_key_commitment(Bytes{}), delta = rc.verifier_key_commitment_delta(keys), parts(2 * 3 * 144, 0), sum(3 * 96);
        for (int k = 0; k < 2; k++) for (int c = 0; c < 3; c++) {
          uint8_t* d = parts.data() + (size_t)(k * 3 + c) * 144;
          const uint8_t* src = k == 0 ? empty.data() + 96 * c : (c < 2 ? delta.data() + 96 * c : nullptr);
          bool inf = true; if (src) for (int j = 0; j < 96; j++) inf = inf && src[j] == 0;
          if (inf) d[48] = 1; else { std::memcpy(d, src, 96); d[96] = 1; }
        }
        eng.check(vrfs_g1_sum_partials(eng.ctx(), 2, 3, parts.data(), sum.data()));
        CHECK(sum == com);
      }
      {   // G1 on the wire: the generator's published compressed encoding, and a round trip of the commitment
        Bytes genc = g1_serialize_compressed(eng, g1);
        CHECK(hex(genc.data(), 48) == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb");
        Bytes blob = rc.ring_commitment_bytes(keys);
        auto [pts, okf] = g1_deserialize_compressed(eng, blob);
        CHECK(blob.size() == 144 && okf[0] == 1 && okf[1] == 1 && okf[2] == 1 && pts == com);
      }
      Bytes ev(cols.size()), back(cols.size());
      eng.check(vrfs_fr_fft_batch(eng.ctx(), 6, 3, 0, cols.data(), ev.data()));
      eng.check(vrfs_fr_fft_batch(eng.ctx(), 6, 3, 1, ev.data(), back.data()));
      CHECK(back == cols && ev != cols);
    }
    // a whole-call failure is an exception, not a verdict
    bool threw = false;
    try { vrfs_status st = vrfs_ietf_verify_batch(eng.ctx(), (vrfs_suite)9, 1, sig.data(), sig.data(), sig.data(), sig.data(), sig.data(), nullptr, nullptr, sig.data(), nullptr); eng.check(st); }
    catch (const CallError& e) { threw = e.status == VRFS_BAD_ARG; }
    CHECK(threw);
  } catch (const CallError& e) {
    std::printf("CallError %d: %s\n", (int)e.status, e.what());
    return 2;
  }
  std::printf(failures ? "FAILED (%d)\n" : "all ok\n", failures);
  return failures ? 1 : 0;
}
