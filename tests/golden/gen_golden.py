"""Regenerates tests/golden/reference_fixtures.json from the reference tree (run in the build
container only: /root/reference does not exist on the GPU box).

Holds (a) the reference's 256-entry FIELD_NORMS_TABLE (src/fieldnorm/code.rs:13) so that the
restated closed form can be checked entry by entry, and (b) the bytes of the on-disk format
fixtures tests/compat_tests_data/index_v{6,7}/*.{idx,fieldnorm,pos,term} that
src/compat_tests.rs:39-56 opens (1 doc each: pins VInt postings, the 8-byte token-count header,
the composite-file footer and the crc/version footer).
"""
import glob
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

src = open(os.path.join(REF, "src/fieldnorm/code.rs")).read()
body = src.split("FIELD_NORMS_TABLE: [u32; 256] = [")[1].split("];")[0]
table = [int(x.replace("_", "")) for x in re.findall(r"[\d_]+", body)]
assert len(table) == 256

fixtures = {}
for ver in ("index_v6", "index_v7"):
    d = os.path.join(REF, "tests/compat_tests_data", ver)
    entry = {"meta": json.load(open(os.path.join(d, "meta.json")))}
    for ext in ("idx", "fieldnorm", "pos", "term"):
        (path,) = glob.glob(os.path.join(d, "*." + ext))
        entry[ext] = open(path, "rb").read().hex()
    fixtures[ver] = entry

json.dump({"field_norms_table": table, "compat": fixtures},
          open(os.path.join(HERE, "reference_fixtures.json"), "w"), indent=1)
print("wrote reference_fixtures.json")
