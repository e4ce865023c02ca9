#!/usr/bin/env python3
"""Extract the reference's own fixtures for this path into tests/golden/*.json.

Run in the build container (needs /root/reference); the JSON files are committed so that tests on the
GPU box, which has no /root/reference, can use them.  Only DATA (hex strings, accept/reject flags) is
extracted -- no reference code.
Sources:
  pairing/bls12381/deserialization_tests/G1/*.yaml, G2/*.yaml   (ZCash-format accept/reject vectors,
      driven in the reference by pairing/bls12381/bls12381_test.go:74-186)
  pairing/bls12381/kilic/suite_test.go:17-72                     (drand signature KATs)
  pairing/bls12381/bls12381_test.go:877-904                      (TestSignatureEdgeCase)
  pairing/bn254/point_test.go:14-124, test_vectors_test.go       (bn254 Keccak/SvdW hash-to-G1 vectors)
  encrypt/ibe/ibe_test.go:202-245                                (the one GT-byte dependent vector; inlined in the tests)
  pairing/bn256/hash_test.go:11-19,45-57                         (HashG1: 11 marshalled points, msg = one byte i, dst = nil)
"""
import glob
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def yaml_vectors(group, key):
    out = []
    for path in sorted(glob.glob(f"{REF}/pairing/bls12381/deserialization_tests/{group}/*.yaml")):
        txt = open(path).read()
        m = re.search(key + r":\s*'?\"?([0-9a-fA-Fx]*)'?\"?\s*}", txt)
        hexstr = m.group(1)
        valid = re.search(r"output:\s*(\S+)", txt).group(1)
        out.append({"name": os.path.basename(path)[:-5], "input": hexstr, "valid": valid != "null"})
    return out


def go_bytes(src, var):
    m = re.search(var + r"\s*:=\s*\[\]byte\{([^}]*)\}", src)
    return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]+)", m.group(1))).hex()


def bn254_hash_vectors():
    pt = open(f"{REF}/pairing/bn254/point_test.go").read()
    tv = open(f"{REF}/pairing/bn254/test_vectors_test.go").read()
    h2p_dst = re.search(r'domain := \[\]byte\("([^"]+)"\)', pt).group(1)
    msg1 = re.search(r'Hash\(\[\]byte\("([^"]+)"\)\)', pt).group(1)
    hexes = re.findall(r'hex\.DecodeString\("([0-9a-f]+)"\)', pt)
    expand = re.search(r'func TestExpandMsg.*?dst := \[\]byte\("([^"]+)"\).*?DecodeString\("([0-9a-f]+)"\).*?!= "([0-9a-f]+)"', pt, re.S)
    f_dst = re.search(r'func TestHashToField.*?dst := \[\]byte\("([^"]+)"\)', pt, re.S).group(1)
    m_dst = re.search(r'func TestMapToPoint.*?dst := \[\]byte\("([^"]+)"\)', pt, re.S).group(1)
    h2f_src, m2p_src = tv.split("var mapToPointTestVectors")
    h2f = [{"msg": m, "x": x, "y": y} for m, x, y in
           re.findall(r'Msg:\s*"([0-9a-f]*)",\s*RefX:\s*"([0-9a-f]+)",\s*RefY:\s*"([0-9a-f]+)"', h2f_src)]
    m2p = [{"u": u, "x": x, "y": y} for u, x, y in   # first 200 of 1000 keep the fixture small
           re.findall(r'U:\s*"(\d+)",\s*RefX:\s*"(\d+)",\s*RefY:\s*"(\d+)"', m2p_src)][:200]
    out = {
        "_source": "pairing/bn254/point_test.go:14-124 + test_vectors_test.go (data only)",
        "hash_to_point": {"dst": h2p_dst, "cases": [{"msg_hex": msg1.encode().hex(), "point": hexes[0]},
                                                    {"msg_hex": hexes[1], "point": hexes[2]}]},
        "expand_msg": {"dst": expand.group(1), "msg_hex": expand.group(2), "len": 96, "out": expand.group(3)},
        "hash_to_field": {"dst": f_dst, "cases": h2f},
        "map_to_point": {"dst": m_dst, "cases": m2p},
    }
    json.dump(out, open(os.path.join(HERE, "bn254_hash_vectors.json"), "w"), indent=1)


def bn256_hashg1_vectors():
    src = open(f"{REF}/pairing/bn256/hash_test.go").read()
    body = src.split("var marshaledHashes")[1]
    rows = re.findall(r"\[64\]byte\{(\d[^}{]*)\}", body)
    out = {"_source": "pairing/bn256/hash_test.go:11-19 (TestKnownHashes: HashG1([]byte{byte(i)}, nil)) and :45-57 (data only)",
           "cases": [{"msg_hex": "%02x" % i, "dst_hex": "", "point": bytes(int(x) for x in re.findall(r"\d+", r)).hex()}
                     for i, r in enumerate(rows)]}
    assert len(out["cases"]) == 11 and all(len(c["point"]) == 128 for c in out["cases"])
    json.dump(out, open(os.path.join(HERE, "bn256_hashg1_vectors.json"), "w"), indent=1)


def main():
    json.dump({"G1": yaml_vectors("G1", "pubkey"), "G2": yaml_vectors("G2", "signature")},
              open(os.path.join(HERE, "bls12381_deserialization.json"), "w"), indent=1)
    st = open(f"{REF}/pairing/bls12381/kilic/suite_test.go").read()
    t = open(f"{REF}/pairing/bls12381/bls12381_test.go").read()
    strs = re.findall(r'(\w+)\s*:=\s*"([0-9a-f]{64,})"', st)
    kat = {
        "sig_on_g1_g2domain": {"pk_g2": strs[0][1], "sig_g1": strs[1][1], "round": 1,
                               "note": "kilic/suite_test.go:17-46: msg = sha256(u64be(round)); verifies ONLY with the G2 DST "
                                       "used for hashing to G1"},
        "sig_on_g2": {"pk_g1": strs[2][1], "sig_g2": strs[3][1], "prev_sig": strs[4][1], "round": 1,
                      "note": "kilic/suite_test.go:48-72: msg = sha256(prev_sig || u64be(round)), default G2 DST"},
        "edge_case_g1": {"pk_g2": go_bytes(t, "publicBytes"), "msg": go_bytes(t, "message"), "sig_g1": go_bytes(t, "sig"),
                         "note": "bls12381_test.go:877-904 TestSignatureEdgeCase: sigs on G1, default G1 DST"},
    }
    json.dump(kat, open(os.path.join(HERE, "bls12381_signature_kats.json"), "w"), indent=1)
    bn254_hash_vectors()
    bn256_hashg1_vectors()
    print("wrote golden fixtures")


if __name__ == "__main__":
    main()
