"""Extracts the expected values of the reference's 72 kernel known-answer tests into one JSON
fixture.  Run in the build container (needs /root/reference); the output is committed because
/root/reference does not exist on the GPU box.

  source : /root/reference/src/gromacs/gmxlib/nonbonded/tests/refdata/
           NBInteraction_NonbondedFepTest_testKernel_{0..71}.xml   (double-precision values)
  output : tests/golden/nb_free_energy_kat.json
"""
import json
import os
import xml.etree.ElementTree as ET

REF = "/root/reference/src/gromacs/gmxlib/nonbonded/tests/refdata"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nb_free_energy_kat.json")


def vec(node):
    return [float(node.find(f"Real[@Name='{c}']").text) for c in "XYZ"]


def main():
    cases = []
    for i in range(72):
        root = ET.parse(os.path.join(REF, f"NBInteraction_NonbondedFepTest_testKernel_{i}.xml")).getroot()
        reals = {r.get("Name").strip(): float(r.text) for r in root.findall("Real")}
        forces = [vec(v) for v in root.find("Sequence[@Name='Forces']").findall("Vector")]
        central = vec(root.find("Shift-Forces").find("Vector[@Name='Central']"))
        cases.append(
            dict(index=i, EVdw=reals["EVdw"], ECoul=reals["ECoul"], dVdlCoul=reals["dVdlCoul"],
                 dVdlVdw=reals["dVdlVdw"], forces=forces, shift_force_central=central)
        )
    with open(OUT, "w") as fh:
        json.dump(dict(source=REF, cases=cases), fh, indent=1)
    print("wrote", OUT, len(cases))


if __name__ == "__main__":
    main()
