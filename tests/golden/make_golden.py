"""Regenerates tests/golden/*.  Run in the build container (needs /root/reference for the assets).

* known_answers.json — vectors that do NOT come from this repo's oracle:
    - the reference's own pinned generator outputs (core/*/src/test_prelude.rs validate_* tests),
    - the one golden byte vector of common/src/transforms/split_565_color_endpoints/tests.rs:129-153,
    - the SURVEY.md §8c known answers (an independent restatement of decorrelate.rs and the
      dispatcher offsets made during the survey).
  They are written out literally below; the oracle is CHECKED against them in tests/test_oracle.py.
    - the headers of the reference's DDS fixtures (src/assets/tests/r2-256-bc{1,2,3,7}.dds),
* r2-256-bc{1,2,3}.payload.zlib — the block payloads (4096 blocks, DDS header stripped) of the
  reference's real-texture fixtures src/assets/tests/r2-256-bc{1,2,3}.dds, zlib-compressed.
"""
import json
import zlib
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src")

KNOWN = {
    "decorrelate": {  # colour -> [var1, var2, var3]   
        "F800": ["BFD1", "5FF1", "BFE2"], "07E0": ["783F", "BC1F", "783F"], "001F": ["F841", "7C21", "F842"],
        "FFFF": ["F820", "FC00", "F801"], "0000": ["0000", "0000", "0000"], "0100": ["1004", "0804", "1008"],
        "0302": ["F79B", "7BDB", "F7B6"], "1234": ["0BAD", "85CD", "0B9B"], "ABCD": ["021E", "011E", "023C"],
    },
    "generators": {  # reference validate_bcN_test_data_generator tests (3 blocks each)
        "bc1": "0001020380818283" "0405060784858687" "08090a0b88898a8b",
        "bc2": "0001020304050607" "80818283c0c1c2c3" "08090a0b0c0d0e0f" "84858687c4c5c6c7"
               "1011121314151617" "88898a8bc8c9cacb",
        "bc3": "0001202122232425" "80818283c0c1c2c3" "0203262728292a2b" "84858687c4c5c6c7"
               "04052c2d2e2f3031" "88898a8bc8c9cacb",
    },
    "split_565": {"in": "000110110405141508091819", "out": "000104050809101114151819"},
    "bc1_3blocks": {  # input = generators.bc1 ; key = "<variant>/<split>"   
        "0/0": "000102030405060708090a0b" "808182838485868788898a8b",
        "0/1": "000104050809" "020306070a0b" "808182838485868788898a8b",
        "1/0": "04109bf7029f89be50e6d705" "808182838485868788898a8b",
        "1/1": "0410029f50e6" "9bf789bed705" "808182838485868788898a8b",
        "2/0": "0408db7b824f495f3073f702" "808182838485868788898a8b",
        "2/1": "0408824f3073" "db7b495ff702" "808182838485868788898a8b",
        "3/0": "0810b6f7049f92be60e6ee05" "808182838485868788898a8b",
        "3/1": "0810049f60e6" "b6f792beee05" "808182838485868788898a8b",
    },
    "bc2_2blocks": {"1/1": "000102030405060708090a0b0c0d0e0f" "1ebc0c83" "855b93a2" "c0c1c2c3c4c5c6c7"},
    "bc3_2blocks": {  # key = "<variant>/<split_alpha>/<split_colour>"
        "1/1/1": "00020103" "202122232425262728292a2b" "1ebc0c83855b93a2" "c0c1c2c3c4c5c6c7",
        "0/0/0": "00010203" "202122232425262728292a2b" "8081828384858687" "c0c1c2c3c4c5c6c7",
        "1/0/0": "00010203" "202122232425262728292a2b" "1ebc855b0c8393a2" "c0c1c2c3c4c5c6c7",
    },
}

if __name__ == "__main__":
    # headers of the reference's DDS fixtures (128 bytes; 148 for the DX10 one) + their file sizes
    KNOWN["dds_fixtures"] = {}
    for name, hdr in (("bc1", 128), ("bc2", 128), ("bc3", 128), ("bc7", 148)):
        dds = (REF / f"assets/tests/r2-256-{name}.dds").read_bytes()
        KNOWN["dds_fixtures"][name] = {"header": dds[:hdr].hex(), "file_len": len(dds)}
    (HERE / "known_answers.json").write_text(json.dumps(KNOWN, indent=1) + "\n")
    for n, bpb in ((1, 8), (2, 16), (3, 16)):
        dds = (REF / f"assets/tests/r2-256-bc{n}.dds").read_bytes()
        payload = dds[128:128 + 4096 * bpb]
        assert len(payload) == 4096 * bpb
        (HERE / f"r2-256-bc{n}.payload.zlib").write_bytes(zlib.compress(payload, 9))
    print("ok")
