"""Extracts the expected values of the reference's perturbed 1-4 pair test into a JSON fixture.

  source : /root/reference/src/gromacs/listed_forces/tests/refdata/
           14Interaction_ListedForcesPairsTest_Ifunc_{0,1,2}.xml
           (LJ14 with c6B = c12B = 0, FEP = Yes; PBC = No / XY / Xyz; lambda 0, 0.5, 1; Beutler, Gapsys;
            inputs in listed_forces/tests/pairs.cpp:160-190, 320-340, 440-460)
  output : tests/golden/pairs14_kat.json
Run in the build container (needs /root/reference); the output is committed.
"""
import json
import os
import xml.etree.ElementTree as ET

REF = "/root/reference/src/gromacs/listed_forces/tests/refdata"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairs14_kat.json")


def vec(node):
    return [float(node.find(f"Real[@Name='{c}']").text) for c in "XYZ"]


def main():
    cases = []
    for idx, pbc in enumerate(["no", "xy", "xyz"]):
        root = ET.parse(os.path.join(REF, f"14Interaction_ListedForcesPairsTest_Ifunc_{idx}.xml")).getroot()
        fep = root.find("FunctionType[@Name='LJ14']").find("FEP[@Name='Yes']")
        for lam in fep.findall("Lambda"):
            for sc in lam.findall("Sofcore"):
                reals = {r.get("Name").strip(): float(r.text) for r in sc.findall("Real")}
                forces = [vec(v) for v in sc.find("Sequence[@Name='Forces']").findall("Vector")]
                central = vec(sc.find("Shift-Forces").find("Vector[@Name='Central']"))
                cases.append(dict(pbc=pbc, lam=float(lam.get("Name")), softcore=sc.get("Name"),
                                  ECoul14=reals["Epot Coulomb14"], ELJ14=reals["Epot LJ14"], dVdlCoul=reals["dVdlCoul"],
                                  dVdlVdw=reals["dVdlVdw"], forces=forces, shift_force_central=central))
    with open(OUT, "w") as fh:
        json.dump(dict(source=REF, cases=cases), fh, indent=1)
    print("wrote", OUT, len(cases))


if __name__ == "__main__":
    main()
