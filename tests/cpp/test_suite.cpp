// Conformance test of the C++ host mirror (kyber_b200/host/kyber_b200.hpp) against the engine, written after
// the reference's own generic tests: testGroup (pairing/bls12381/bls12381_test.go:196-418 = util/test/test.go:
// 325-401), the pairing property tests (:448-474, :580-631) and sign/bls tests (sign/bls/bls_test.go).
// Run by tests/test_gpu_cpp_host.py on the GPU box.  Prints "ok <name>" per check; exit code != 0 on failure.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../kyber_b200/host/kyber_b200.hpp"

using namespace b200;
static int fails = 0;
#define CHECK(name, cond) do { if (cond) printf("ok %s\n", name); else { printf("FAIL %s (line %d)\n", name, __LINE__); fails++; } } while (0)

static Scalar pick(std::mt19937_64& g) {       // Scalar.Pick stand-in: uniform bytes reduced mod r
  Bytes b(40);
  for (auto& x : b) x = (uint8_t)g();
  Scalar s; s.SetBytes(b); return s;
}

template <class E, class G>
static void test_group(const char* tag, const G& grp, std::shared_ptr<Engine> eng, std::mt19937_64& rng) {
  char nm[128];
  auto N = [&](const char* s) { snprintf(nm, sizeof nm, "%s.%s", tag, s); return nm; };
  Scalar s1 = pick(rng), s2 = pick(rng), st;
  E gen(eng), p1(eng), p2(eng), dh1(eng), dh2(eng), pt(eng), zero(eng);
  gen.Base(); zero.Null();
  CHECK(N("generator_not_identity"), !gen.Equal(zero));
  // Diffie-Hellman: s2*(s1*G) == s1*(s2*G)
  p1.Mul(s1, nullptr); p2.Mul(s2, nullptr);
  CHECK(N("pubkeys_differ"), !p1.Equal(p2));
  dh1.Mul(s2, &p1); dh2.Mul(s1, &p2);
  CHECK(N("dh"), dh1.Equal(dh2));
  // additive homomorphism: (s1+s2)G == s1 G + s2 G ; Sub ; Neg
  st.Add(s1, s2); pt.Mul(st, nullptr);
  E sum(eng); sum.Add(p1, p2);
  CHECK(N("add_homomorphism"), pt.Equal(sum));
  E back(eng); back.Sub(sum, p2);
  CHECK(N("sub"), back.Equal(p1));
  E neg(eng); neg.Neg(p1); E z(eng); z.Add(p1, neg);
  CHECK(N("neg_gives_identity"), z.Equal(zero));
  // multiplicative: s1*(s2*G) == (s1*s2)*G ; division
  st.Mul(s1, s2); pt.Mul(st, nullptr);
  CHECK(N("mul_homomorphism"), pt.Equal(dh1));
  Scalar inv; inv.Div(st, s2);
  CHECK(N("scalar_div"), inv.Equal(s1));
  // aliasing receiver == argument (bls.go:73)
  E al(eng); al.Set(p1); al.Mul(s2, &al);
  CHECK(N("aliasing_mul"), al.Equal(dh1));
  al.Set(p1); al.Add(al, al); E dbl(eng); Scalar two(2); dbl.Mul(two, &p1);
  CHECK(N("aliasing_add_doubling"), al.Equal(dbl));
  // marshal round trip, identity included; wrong sizes and garbage are errors
  Bytes enc = p1.MarshalBinary();
  CHECK(N("marshal_size"), (int)enc.size() == grp.PointLen());
  E dec(eng); dec.UnmarshalBinary(enc);
  CHECK(N("marshal_roundtrip"), dec.Equal(p1));
  Bytes ze = zero.MarshalBinary();
  CHECK(N("identity_encoding"), ze[0] == 0xC0);
  dec.UnmarshalBinary(ze);
  CHECK(N("identity_roundtrip"), dec.Equal(zero));
  bool threw = false;
  try { Bytes bad(enc.begin(), enc.end() - 1); dec.UnmarshalBinary(bad); } catch (const std::runtime_error&) { threw = true; }
  CHECK(N("unmarshal_wrong_size_is_error"), threw);
  threw = false;
  try { Bytes bad = enc; bad[0] &= 0x7f; dec.UnmarshalBinary(bad); } catch (const std::runtime_error&) { threw = true; }
  CHECK(N("unmarshal_bad_flags_is_error"), threw);
  CHECK(N("is_in_correct_group"), p1.IsInCorrectGroup());
  // batch extension == loops
  std::vector<Scalar> ss; std::vector<E> pp;
  E acc(eng); acc.Null();
  for (int i = 0; i < 9; i++) { ss.push_back(pick(rng)); E q(eng); q.Mul(pick(rng), nullptr); pp.push_back(q); }
  auto mb = grp.MulBatch(ss, pp);
  bool all = true;
  for (int i = 0; i < 9; i++) { E t(eng); t.Mul(ss[i], &pp[i]); all = all && t.Equal(mb[i]); acc.Add(acc, t); }
  CHECK(N("mul_batch_equals_loop"), all);
  CHECK(N("msm_equals_mul_add_loop"), grp.MSM(ss, pp).Equal(acc));
}

int main() {
  std::mt19937_64 rng(20260923);
  Suite suite(0);
  auto eng = suite.engine();
  // scalar wire format (TestScalarEndianess, bls12381_test.go:41-72)
  Scalar one; one.One();
  Bytes ob = one.MarshalBinary();
  CHECK("scalar.big_endian_one", ob.size() == 32 && ob[31] == 1 && ob[0] == 0);
  Scalar a = pick(rng), ai, prod; ai.Inv(a); prod.Mul(a, ai);
  CHECK("scalar.inverse", prod.Equal(one));
  bool threw = false;
  try { Bytes rb(32); for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) rb[(3 - i) * 8 + k] = (uint8_t)(Scalar::R[i] >> (56 - 8 * k)); Scalar t; t.UnmarshalBinary(rb); } catch (const std::runtime_error&) { threw = true; }
  CHECK("scalar.unmarshal_rejects_modulus", threw);
  CHECK("group.names", suite.G1().String() == "bls12-381.G1" && suite.G2().String() == "bls12-381.G2" && suite.G1().PointLen() == 48 && suite.G2().PointLen() == 96 && suite.G1().ScalarLen() == 32);

  test_group<G1Elt>("G1", suite.G1(), eng, rng);
  test_group<G2Elt>("G2", suite.G2(), eng, rng);

  // pairing: bilinearity e(aG1,bG2) == e(abG1,G2) (bls12381_test.go:448-474) and the a*b = c+d identity via
  // ValidatePairing semantics e(p1,p2) == e(inv1,inv2)
  Scalar sa = pick(rng), sb = pick(rng), sab; sab.Mul(sa, sb);
  G1Elt pa(eng), pab(eng), g1(eng); G2Elt qb(eng), g2(eng);
  g1.Base(); g2.Base(); pa.Mul(sa, nullptr); qb.Mul(sb, nullptr); pab.Mul(sab, nullptr);
  GTElt e1 = suite.Pair(pa, qb), e2 = suite.Pair(pab, g2);
  CHECK("pairing.bilinear_gt_equal", e1.Equal(e2));
  CHECK("pairing.gt_size", e1.MarshalSize() == 576 && !e1.Equal(suite.Pair(g1, g2)));
  CHECK("pairing.validate_true", suite.ValidatePairing(pa, qb, pab, g2));
  G1Elt wrong(eng); Scalar sab1; sab1.Add(sab, one); wrong.Mul(sab1, nullptr);
  CHECK("pairing.validate_false", !suite.ValidatePairing(pa, qb, wrong, g2));
  auto vb = suite.ValidatePairingBatch({pa, pa}, {qb, qb}, {pab, wrong}, {g2, g2});
  CHECK("pairing.validate_batch", vb.size() == 2 && vb[0] && !vb[1]);
  bool panicked = false;
  try { suite.Pair(qb, pa); } catch (const std::logic_error&) { panicked = true; }
  CHECK("pairing.wrong_group_panics", panicked);

  // GT as a kyber.Group (kilic/gt.go:33-83): Add = product, Neg = inverse, Mul = exponentiation, Null = 1
  {
    CHECK("gt.group_names", suite.GT().String() == "bls12-381.GT" && suite.GT().PointLen() == 576 && !suite.GT().IsPrimeOrder());
    GTElt ea = suite.Pair(pa, g2), eb = suite.Pair(g1, qb), one = suite.GT().NewPoint();
    GTElt sum = suite.GT().NewPoint(); sum.Add(ea, eb);                    // e(aG1, G2) e(G1, bG2) = e(G1, G2)^(a+b)
    Scalar apb; apb.Add(sa, sb);
    GTElt base = suite.Pair(g1, g2), pw = suite.GT().NewPoint(); pw.Mul(apb, base);
    CHECK("gt.add_is_product_mul_is_exp", sum.Equal(pw));
    GTElt ng = suite.GT().NewPoint(); ng.Neg(ea);
    GTElt z = suite.GT().NewPoint(); z.Add(ea, ng);
    CHECK("gt.neg_gives_null", z.Equal(one));
    GTElt df = suite.GT().NewPoint(); df.Sub(sum, eb);
    CHECK("gt.sub", df.Equal(ea));
    GTElt e12 = suite.GT().NewPoint(); e12.Mul(sb, ea);                     // e(aG1, G2)^b == e(aG1, bG2)
    CHECK("gt.mul_bilinear", e12.Equal(e1));
    GTElt al = ea.Clone(); al.Add(al, al); Scalar two(2); GTElt dbl = suite.GT().NewPoint(); dbl.Mul(two, ea);
    CHECK("gt.aliasing_add", al.Equal(dbl));
    GTElt rt = suite.GT().NewPoint(); rt.UnmarshalBinary(ea.MarshalBinary());
    CHECK("gt.marshal_roundtrip", rt.Equal(ea) && ea.MarshalSize() == 576);
    bool bad = false;
    try { Bytes b(576, 0xff); rt.UnmarshalBinary(b); } catch (const std::runtime_error&) { bad = true; }
    CHECK("gt.unmarshal_rejects_noncanonical", bad);
    bool pan = false;
    try { one.Base(); } catch (const std::logic_error&) { pan = true; }
    CHECK("gt.base_panics", pan);
    // Miller / Finalize and the n-pair product with one final exponentiation
    CHECK("pairing.finalize_of_miller_is_pair", suite.Finalize(suite.Miller(pa, qb)).Equal(e1));
    GTElt m2 = suite.GT().NewPoint(); m2.Add(suite.Miller(pa, g2), suite.Miller(g1, qb));
    CHECK("pairing.one_final_exp_for_a_product", suite.Finalize(m2).Equal(sum));
    G1Elt npab(eng); npab.Neg(pab);
    CHECK("pairing.product_is_one", suite.PairingProductIsOne({pa, npab}, {qb, g2}));
    CHECK("pairing.product_is_not_one", !suite.PairingProductIsOne({pa, pab}, {qb, g2}));
  }

  // group/edwards25519 through the engine: s*B chains and the a*(b*B) == b*(a*B) exchange of examples/dh_test.go
  {
    ed25519::Curve ed(eng);
    ed25519::PointBytes sa2{}, sb2{};
    for (int i = 0; i < 31; i++) { sa2[i] = (uint8_t)rng(); sb2[i] = (uint8_t)rng(); }
    sa2[31] = 0x0f; sb2[31] = 0x0e;                                           // < 2^253: inside the a[31] <= 127 precondition
    ed25519::PointBytes A = ed.Mul(sa2, nullptr), B = ed.Mul(sb2, nullptr);
    CHECK("ed25519.dh", ed.Mul(sa2, &B) == ed.Mul(sb2, &A));
    ed25519::PointBytes onele{}; onele[0] = 1;
    CHECK("ed25519.one_times_base", ed.Mul(onele, nullptr) == ed25519::BASE);
    auto outs = ed.MulBatch({sa2, sb2, onele}, {ed25519::BASE, ed25519::BASE, A});
    CHECK("ed25519.batch", outs[0] == A && outs[1] == B && outs[2] == A);
  }

  // sign/bls on G1 (sign/bls/bls.go:33-96): sign, verify, reject wrong message / key / mangled signature
  SchemeOnG1 scheme(suite);
  Scalar sk = pick(rng), sk2 = pick(rng);
  G2Elt pk(eng), pk2(eng); pk.Mul(sk, nullptr); pk2.Mul(sk2, nullptr);
  Bytes msg = {'H', 'e', 'l', 'l', 'o', ' ', 'B', 'o', 'n', 'e', 'h', '-', 'L', 'y', 'n', 'n', '-', 'S', 'h', 'a', 'c', 'h', 'a', 'm'};
  Bytes sig = scheme.Sign(sk, msg);
  CHECK("bls.sign_size", sig.size() == 48);
  CHECK("bls.verify", scheme.Verify(pk, msg, sig));
  Bytes msg2 = msg; msg2[0] ^= 1;
  CHECK("bls.verify_wrong_msg_fails", !scheme.Verify(pk, msg2, sig));
  CHECK("bls.verify_wrong_key_fails", !scheme.Verify(pk2, msg, sig));
  Bytes sig2 = sig; sig2[20] ^= 0x40;
  CHECK("bls.verify_mangled_sig_fails", !scheme.Verify(pk, msg, sig2));
  // TestSignatureEdgeCase (bls12381_test.go:877-904), bytes from the reference test
  static const uint8_t pkb[96] = {0x83, 0xcf, 0xf, 0x28, 0x96, 0xad, 0xee, 0x7e, 0xb8, 0xb5, 0xf0, 0x1f, 0xca, 0xd3, 0x91, 0x22, 0x12, 0xc4, 0x37, 0xe0, 0x7, 0x3e, 0x91, 0x1f, 0xb9, 0x0, 0x22, 0xd3, 0xe7, 0x60, 0x18, 0x3c, 0x8c, 0x4b, 0x45, 0xb, 0x6a, 0xa, 0x6c, 0x3a, 0xc6, 0xa5, 0x77, 0x6a, 0x2d, 0x10, 0x64, 0x51, 0xd, 0x1f, 0xec, 0x75, 0x8c, 0x92, 0x1c, 0xc2, 0x2b, 0xe, 0x17, 0xe6, 0x3a, 0xaf, 0x4b, 0xcb, 0x5e, 0xd6, 0x63, 0x4, 0xde, 0x9c, 0xf8, 0x9, 0xbd, 0x27, 0x4c, 0xa7, 0x3b, 0xab, 0x4a, 0xf5, 0xa6, 0xe9, 0xc7, 0x6a, 0x4b, 0xc0, 0x9e, 0x76, 0xea, 0xe8, 0x99, 0x1e, 0xf5, 0xec, 0xe4, 0x5a};
  static const uint8_t mb[32] = {0xa1, 0xc6, 0xbe, 0xc3, 0xb9, 0xa6, 0xf0, 0x98, 0x9d, 0x4d, 0x80, 0x2d, 0xbf, 0xe2, 0xb9, 0xb, 0x49, 0x5f, 0xa1, 0x74, 0x2b, 0x58, 0x99, 0x63, 0x45, 0x1e, 0xeb, 0xa9, 0xb1, 0x87, 0xb8, 0x15};
  static const uint8_t sgb[48] = {0x95, 0x89, 0x0, 0x9b, 0x47, 0xbf, 0xd9, 0xe3, 0x65, 0x10, 0x6b, 0x11, 0xa3, 0x42, 0xfe, 0x50, 0x75, 0xeb, 0x44, 0x5, 0xb0, 0x2b, 0x80, 0xe8, 0x93, 0x42, 0x69, 0x86, 0xcf, 0xb6, 0x0, 0x77, 0x99, 0x8e, 0x3b, 0x47, 0x99, 0x68, 0x86, 0xe0, 0x35, 0xca, 0x1c, 0xde, 0x5f, 0xd9, 0x62, 0x89};
  G2Elt epk(eng); epk.UnmarshalBinary(Bytes(pkb, pkb + 96));
  CHECK("bls.TestSignatureEdgeCase", scheme.Verify(epk, Bytes(mb, mb + 32), Bytes(sgb, sgb + 48)));

  // sign/bdn (sign/bdn/bdn_test.go: TestBDN_AggregateSignatures, _SubsetSignature, _RogueAttack shape): aggregate
  // signature of a subset verifies under the aggregate key of the same subset and under no other; the engine's MSM
  // equals the reference's Mul+Add loop with the coefficients c_i + 1.
  {
    bdn::SchemeOnG1 bs(suite);
    const int n = 5;
    std::vector<Scalar> sks; std::vector<G2Elt> pks; std::vector<Bytes> all_sigs;
    for (int i = 0; i < n; i++) { sks.push_back(pick(rng)); G2Elt k(eng); k.Mul(sks[i], nullptr); pks.push_back(k); all_sigs.push_back(bs.Sign(sks[i], msg)); }
    bdn::Mask mask(suite, pks);
    CHECK("bdn.mask_len", mask.Len() == 1 && mask.CountEnabled() == 0 && mask.CountTotal() == n);
    mask.SetBit(0, true); mask.SetBit(2, true); mask.SetBit(4, true);
    std::vector<Bytes> sub = {all_sigs[0], all_sigs[2], all_sigs[4]};
    G1Elt asig = bs.AggregateSignatures(sub, mask);
    G2Elt akey = bs.AggregatePublicKeys(mask);
    CHECK("bdn.aggregate_verifies", bs.Verify(akey, msg, asig.MarshalBinary()));
    // loop form of the reference: sum (c_i * S_i + S_i)
    G1Elt loop(eng); loop.Null();
    for (int i : {0, 2, 4}) { G1Elt sgi(eng); sgi.UnmarshalBinary(all_sigs[i]); G1Elt t(eng); t.Mul(mask.Factors()[i], &sgi); loop.Add(loop, t); }
    CHECK("bdn.msm_equals_reference_loop", loop.Equal(asig));
    bdn::Mask other(suite, pks); other.SetBit(0, true); other.SetBit(1, true); other.SetBit(4, true);
    CHECK("bdn.wrong_subset_key_fails", !bs.Verify(bs.AggregatePublicKeys(other), msg, asig.MarshalBinary()));
    bool err = false;
    try { bs.AggregateSignatures({all_sigs[0]}, mask); } catch (const std::runtime_error&) { err = true; }
    CHECK("bdn.signature_count_mismatch_is_error", err);
    err = false;
    try { mask.SetBit(n, true); } catch (const std::runtime_error&) { err = true; }
    CHECK("bdn.mask_index_out_of_range_is_error", err);
    // plain sum of keys (no coefficients) must NOT verify the BDN aggregate: the coefficients are in effect
    G2Elt plain(eng); plain.Null(); for (int i : {0, 2, 4}) plain.Add(plain, pks[i]);
    CHECK("bdn.plain_key_sum_fails", !bs.Verify(plain, msg, asig.MarshalBinary()));
  }

  // share (share/poly_test.go shape: TestPublicCheck, TestPublicRecovery): commitments of a random polynomial, every private
  // share checks against them, a wrong share does not, t public shares recover the commitment of the secret
  {
    const int t = 4, n = 6;
    std::vector<Scalar> coef; std::vector<G1Elt> commits;
    for (int j = 0; j < t; j++) { coef.push_back(pick(rng)); G1Elt c(eng); c.Mul(coef[j], nullptr); commits.push_back(c); }
    share::PubPoly pub(suite, commits);
    auto eval = [&](uint32_t i) { Scalar x, acc; x.SetInt64((int64_t)i + 1); acc.Zero(); for (int j = t - 1; j >= 0; j--) { acc.Mul(acc, x); acc.Add(acc, coef[j]); } return acc; };
    std::vector<share::PriShare> pri;
    for (uint32_t i = 0; i < (uint32_t)n; i++) pri.push_back({i, eval(i)});
    bool all = true;
    for (auto& sh : pri) all = all && pub.Check(sh);
    CHECK("share.check_every_share", all);
    share::PriShare bad = pri[2]; bad.V.Add(bad.V, one);
    CHECK("share.check_rejects_wrong_share", !pub.Check(bad));
    G1Elt v3(eng); v3.Mul(pri[3].V, nullptr);
    CHECK("share.eval_equals_share_times_base", pub.Eval(3).V.Equal(v3));
    std::vector<G1Elt> commits2 = commits; commits2[1].Mul(pick(rng), nullptr);
    share::PubPoly pub2(suite, commits2);
    auto grid = share::PubPoly::CheckBatch({&pub, &pub2}, {pri, pri});
    bool row0 = true, row1 = false;
    for (int k = 0; k < n; k++) { row0 = row0 && grid[0][k]; row1 = row1 || grid[1][k]; }
    CHECK("share.check_batch_two_dealers", row0 && !row1);
    std::vector<share::PubShare> pubs = pub.EvalBatch({5, 1, 2, 4});
    std::sort(pubs.begin(), pubs.end(), [](const share::PubShare& a, const share::PubShare& b) { return a.I < b.I; });
    CHECK("share.recover_commit", share::RecoverCommit(suite, pubs).Equal(pub.Commit()));
  }

  printf("%s: %d failure(s)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
