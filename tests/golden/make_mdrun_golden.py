#!/usr/bin/env python
"""tests/golden/mdrun_fe_refdata.json from the reference's own golden vectors of its mdrun free-energy test
(src/programs/mdrun/tests/refdata/EquivalentToReference_FreeEnergyReferenceTest_WithinTolerances_<name>_s.xml,
mixed precision; src/programs/mdrun/tests/freeenergy.cpp).  Needs /root/reference; the JSON is committed."""
import json
import os
import re
import xml.etree.ElementTree as ET

REF = "/root/reference/src/programs/mdrun/tests/refdata"
HERE = os.path.dirname(os.path.abspath(__file__))
SYSTEMS = ["coulandvdwsequential_coul", "coulandvdwsequential_vdw", "coulandvdwtogether", "expanded", "relative",
           "relative-position-restraints", "transformAtoB", "vdwalone"]

out = {}
for name in SYSTEMS:
    path = os.path.join(REF, f"EquivalentToReference_FreeEnergyReferenceTest_WithinTolerances_{name.replace('-', '_')}_s.xml")
    root = ET.parse(path).getroot()
    terms = {}
    for energy in root.iter("Energy"):
        steps, values = [], []
        for real in energy.iter("Real"):
            steps.append(int(re.search(r"Step (\d+)", real.get("Name")).group(1)))
            values.append(float(real.text))
        terms[energy.get("Name")] = dict(steps=steps, values=values)
    out[name] = terms
with open(os.path.join(HERE, "mdrun_fe_refdata.json"), "w") as fh:
    json.dump(dict(source="reference src/programs/mdrun/tests/refdata/*_s.xml (mixed precision)", systems=out), fh, indent=0)
print({k: {t: len(v["values"]) for t, v in d.items()} for k, d in out.items()})
