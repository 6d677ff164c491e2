"""Extract the literal golden vectors of the reference's own tests into JSON fixtures.

Run in the build container (where /root/reference is mounted):
    python tests/golden/extract_goldens.py
It scans the listed Julia test files for bracketed lists that contain only numeric
literals (>= 2 numbers) and records them in file order together with the line on
which each list starts.  No reference *code* is copied -- only the numbers the
reference's tests compare against, which is exactly what pins the oracle
(SURVEY.md section 8c).  /root/reference does not exist on the GPU box, so the tests read
the committed JSON, never the Julia files.
"""
import json
import os
import re
import sys

REF = os.environ.get("GEMPIC_REFERENCE", "/root/reference")
FILES = [
    "test/test_particle_mesh_coupling_spline_1d.jl",
    "test/test_particle_mesh_coupling_spline_2d.jl",
    "test/test_hamiltonian_splitting.jl",
    "test/test_hamiltonian_splitting_boris.jl",
]
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"
LIST = re.compile(r"\[\s*(" + NUM + r"(?:\s*[,;\s]\s*" + NUM + r")+)\s*,?\s*\]", re.S)


def extract(path):
    text = open(path).read()
    out = []
    for m in LIST.finditer(text):
        line = text.count("\n", 0, m.start()) + 1
        vals = [float(v) for v in re.findall(NUM, m.group(1))]
        out.append({"line": line, "values": vals})
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for rel in FILES:
        src = os.path.join(REF, rel)
        data = {"source": rel, "lists": extract(src)}
        name = os.path.splitext(os.path.basename(rel))[0] + ".json"
        with open(os.path.join(here, name), "w") as f:
            json.dump(data, f, indent=1)
        print(name, [(l["line"], len(l["values"])) for l in data["lists"]])


if __name__ == "__main__":
    sys.exit(main())
