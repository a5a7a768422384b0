#!/usr/bin/env python
"""Run inputs for BASELINE.json's configs[0] and configs[1] as real GROMACS systems.

  c1_methane     methane decoupling in a 3.0 nm TIP3P box (~2.7 k atoms), Beutler soft-core
                 (sc-alpha 0.5), PME real space, one lambda (no foreign states)
  c3_hexadecane  the solute of c2_hexadecane in a 10 nm box (~100 k atoms), same settings (made on request only)
  c2_hexadecane  a 50-atom solute (hexadecane, C16H34) transformed A -> B (hydrogens vanish, carbons
                 become united atoms with half the charge) in a 6.3 nm TIP3P box (~25 k atoms), PME,
                 20 lambda states with foreign-energy output

Made with the reference's own tools (gmx solvate / grompp / mdrun of integration/_gmx, force field
files of /root/reference/share/top): solvate, steepest-descent relaxation on the reference's CPU
path, then the 20-step MD run input that tests/test_mdrun_dropin.py runs twice (reference CPU
kernel vs libfepb200).  Output: tests/golden/mdrun_tpr/<name>.tpr.  Needs /root/reference, so it
runs in the build container only; the .tpr files are committed.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
GMX = os.path.join(ROOT, "integration", "_gmx", "bin", "gmx")
ENV = dict(os.environ, GMXLIB="/root/reference/share/top",
           LD_LIBRARY_PATH=os.path.join(ROOT, "integration", "_gmx", "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
ENV.pop("GMX_FEPB200", None)
OUT = os.path.join(ROOT, "tests", "golden", "mdrun_tpr")
WORK = "/tmp/fepb200_systems"

R_CH, R_CC = 0.109, 0.1529


def gmx(args, cwd, stdin=None):
    r = subprocess.run([GMX, "-quiet"] + args, cwd=cwd, env=ENV, input=stdin, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
        raise SystemExit(f"gmx {' '.join(args[:1])} failed")
    return r


def write_gro(path, title, resname, names, xyz, box):
    with open(path, "w") as fh:
        fh.write(f"{title}\n{len(names):5d}\n")
        for i, (n, p) in enumerate(zip(names, xyz), start=1):
            fh.write(f"{1:5d}{resname:<5s}{n:>5s}{i:5d}{p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}\n")
        fh.write(f"{box:10.5f}{box:10.5f}{box:10.5f}\n")


def unit(v):
    return v / np.linalg.norm(v)


# ---------------------------------------------------------------------------------------------
# solutes
# ---------------------------------------------------------------------------------------------
def methane(box):
    c = np.full(3, box / 2)
    t = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], float) / np.sqrt(3.0)
    names = ["C1", "H1", "H2", "H3", "H4"]
    xyz = np.vstack([c, c + R_CH * t])
    itp = """[ moleculetype ]
; name  nrexcl
methane  3

[ atoms ]
;  nr  type      resnr  res  atom  cgnr  charge     mass
   1   opls_138  1      MET  C1    1     -0.240    12.011
   2   opls_140  1      MET  H1    1      0.060     1.008
   3   opls_140  1      MET  H2    1      0.060     1.008
   4   opls_140  1      MET  H3    1      0.060     1.008
   5   opls_140  1      MET  H4    1      0.060     1.008

[ bonds ]
1 2 1
1 3 1
1 4 1
1 5 1

[ angles ]
2 1 3 1
2 1 4 1
2 1 5 1
3 1 4 1
3 1 5 1
4 1 5 1
"""
    return "MET", names, xyz, itp, "methane"


def hexadecane(box):
    """All-trans C16H34 along x through the box centre; state B: hydrogens without charge and LJ,
    carbons as united atoms (sigma 0.3905 nm, epsilon 0.4937 kJ/mol) with half the charge."""
    n_c = 16
    half = np.deg2rad(112.7) / 2
    dx, dy = R_CC * np.sin(half), R_CC * np.cos(half)
    carbons = np.array([[i * dx, (i % 2) * dy, 0.0] for i in range(n_c)])
    carbons += np.full(3, box / 2) - carbons.mean(axis=0)
    names, xyz, bonds, heavy_of = [], [], [], []
    index_of_c = []
    ez = np.array([0.0, 0.0, 1.0])
    for i in range(n_c):
        ci = carbons[i]
        index_of_c.append(len(names))
        names.append(f"C{i + 1}")
        xyz.append(ci)
        hs = []
        if 0 < i < n_c - 1:
            u = unit((ci - carbons[i - 1]) + (ci - carbons[i + 1]))
            a = np.deg2rad(107.8) / 2
            hs = [ci + R_CH * (np.cos(a) * u + s * np.sin(a) * ez) for s in (1, -1)]
        else:
            nb = carbons[1] if i == 0 else carbons[n_c - 2]
            d = unit(ci - nb)
            e1 = unit(np.cross(ez, d))
            a = np.deg2rad(180.0 - 109.5)
            for phi in np.deg2rad([0.0, 120.0, 240.0]):
                hs.append(ci + R_CH * (np.cos(a) * d + np.sin(a) * (np.cos(phi) * e1 + np.sin(phi) * ez)))
        for k, h in enumerate(hs):
            names.append(f"H{i + 1}{'ABC'[k]}")
            xyz.append(h)
            bonds.append((index_of_c[i], len(names) - 1))
    for i in range(n_c - 1):
        bonds.append((index_of_c[i], index_of_c[i + 1]))
    n = len(names)
    assert n == 50
    adj = [[] for _ in range(n)]
    for a, b in bonds:
        adj[a].append(b)
        adj[b].append(a)
    angles = sorted({(min(a, c), b, max(a, c)) for b in range(n) for a in adj[b] for c in adj[b] if a != c})
    dihedrals, pairs = set(), set()
    for b, c in bonds + [(y, x) for x, y in bonds]:
        for a in adj[b]:
            for d in adj[c]:
                if a != c and d != b and a != d:
                    t = (a, b, c, d)
                    if t[::-1] not in dihedrals:
                        dihedrals.add(t)
                    pairs.add((min(a, d), max(a, d)))
    lines = ["[ moleculetype ]", "; name  nrexcl", "hexadecane  3", "", "[ atoms ]",
             ";  nr  type  resnr  res  atom  cgnr  charge  mass  typeB  chargeB  massB"]
    for i, nm in enumerate(names):
        if nm.startswith("C"):
            terminal = len([j for j in adj[i] if names[j].startswith("H")]) == 3
            ta, qa = ("opls_135", -0.18) if terminal else ("opls_136", -0.12)
            lines.append(f"{i + 1:5d} {ta} 1 HEX {nm:5s} {i + 1:3d} {qa:8.3f} 12.011  fep_CU {qa / 2:8.3f} 12.011")
        else:
            lines.append(f"{i + 1:5d} opls_140 1 HEX {nm:5s} {i + 1:3d} {0.06:8.3f}  1.008  fep_HD {0.0:8.3f}  1.008")
    # the charges of state B must add up to the charge of state A for PME: spread the rest over the carbons
    q_b = sum((-0.09 if len([j for j in adj[i] if names[j].startswith('H')]) == 3 else -0.06)
              for i, nm in enumerate(names) if nm.startswith("C"))
    fix = -q_b / n_c
    out = []
    for ln in lines:
        parts = ln.split()
        if len(parts) == 11 and parts[8] == "fep_CU":
            parts[9] = f"{float(parts[9]) + fix:.4f}"
            ln = " ".join(parts)
        out.append(ln)
    lines = out
    lines += ["", "[ bonds ]"] + [f"{a + 1} {b + 1} 1" for a, b in sorted(bonds)]
    lines += ["", "[ pairs ]"] + [f"{a + 1} {b + 1} 1" for a, b in sorted(pairs)]
    lines += ["", "[ angles ]"] + [f"{a + 1} {b + 1} {c + 1} 1" for a, b, c in angles]
    lines += ["", "[ dihedrals ]"] + [f"{a + 1} {b + 1} {c + 1} {d + 1} 3" for a, b, c, d in sorted(dihedrals)]
    return "HEX", names, np.array(xyz), "\n".join(lines) + "\n", "hexadecane"


# ---------------------------------------------------------------------------------------------
# run parameters
# ---------------------------------------------------------------------------------------------
COMMON = """
cutoff-scheme = Verlet
nstlist = 10
pbc = xyz
verlet-buffer-tolerance = 0.005
coulombtype = PME
rcoulomb = 1.0
vdwtype = cut-off
vdw-modifier = potential-shift
rvdw = 1.0
fourierspacing = 0.12
pme-order = 4
ewald-rtol = 1e-5
constraints = h-bonds
"""

EM = "integrator = steep\nnsteps = 300\nemtol = 200\nemstep = 0.01\n" + COMMON

MD = """
integrator = md
dt = 0.002
nsteps = 20
comm-mode = Linear
nstcomm = 10
nstxout = 0
nstvout = 0
nstfout = 0
nstlog = 20
nstcalcenergy = 1
nstenergy = 1
tcoupl = v-rescale
tc-grps = system
tau-t = 0.5
ref-t = 298
pcoupl = no
gen-vel = yes
gen-temp = 298
gen-seed = 20261018
ld-seed = 20261018
""" + COMMON

FEP_C1 = """
free-energy = yes
couple-moltype = methane
couple-lambda0 = vdw-q
couple-lambda1 = none
couple-intramol = no
init-lambda = 0.5
nstdhdl = 5
sc-alpha = 0.5
sc-power = 1
sc-r-power = 6
sc-sigma = 0.3
sc-coul = no
separate-dhdl-file = no
dhdl-derivatives = yes
"""

_LAM = " ".join(f"{v:.4f}" for v in np.linspace(0.0, 1.0, 20))
FEP_C2 = f"""
free-energy = yes
init-lambda-state = 10
fep-lambdas = {_LAM}
calc-lambda-neighbors = -1
nstdhdl = 5
sc-alpha = 0.5
sc-power = 1
sc-r-power = 6
sc-sigma = 0.3
sc-coul = no
separate-dhdl-file = no
dhdl-derivatives = yes
"""

EXTRA_TYPES = """
[ atomtypes ]
; state-B types of the transformed solute
 fep_HD  HC  1   1.00800  0.000  A  0.00000e+00  0.00000e+00
 fep_CU  CT  6  12.01100  0.000  A  3.90500e-01  4.93712e-01
"""


def mdp_with(base, overrides):
    """`base` with the keys of `overrides` replaced (or appended)."""
    keys = {k.strip().lower().replace("_", "-") for k in overrides}
    kept = [ln for ln in base.splitlines()
            if "=" not in ln or ln.split("=")[0].strip().lower().replace("_", "-") not in keys]
    return "\n".join(kept + [f"{k} = {v}" for k, v in overrides.items()]) + "\n"


def build(name, box, solute, fep, variants=None):
    work = os.path.join(WORK, name)
    os.makedirs(work, exist_ok=True)
    resname, names, xyz, itp, molname = solute(box)
    write_gro(os.path.join(work, "solute.gro"), name, resname, names, xyz, box)
    with open(os.path.join(work, "topol.top"), "w") as fh:
        fh.write('#include "oplsaa.ff/forcefield.itp"\n' + EXTRA_TYPES + "\n" + itp + '\n#include "oplsaa.ff/tip3p.itp"\n\n'
                 f"[ system ]\n{name}\n\n[ molecules ]\n{molname} 1\n")
    gmx(["solvate", "-cp", "solute.gro", "-cs", "spc216.gro", "-o", "solvated.gro", "-p", "topol.top"], work)
    for stage, mdp in (("em", EM), ("md", MD + fep)):
        with open(os.path.join(work, stage + ".mdp"), "w") as fh:
            fh.write(mdp)
    gmx(["grompp", "-f", "em.mdp", "-c", "solvated.gro", "-p", "topol.top", "-o", "em.tpr", "-maxwarn", "10"], work)
    gmx(["mdrun", "-s", "em.tpr", "-deffnm", "em", "-nb", "cpu", "-pme", "cpu", "-bonded", "cpu", "-ntmpi", "1",
         "-ntomp", "8"], work)
    gmx(["grompp", "-f", "md.mdp", "-c", "em.gro", "-p", "topol.top", "-o", "md.tpr", "-maxwarn", "10"], work)
    os.makedirs(OUT, exist_ok=True)
    dst = os.path.join(OUT, name + ".tpr")
    with open(os.path.join(work, "md.tpr"), "rb") as src, open(dst, "wb") as out:
        out.write(src.read())
    natoms = int(open(os.path.join(work, "em.gro")).read().split("\n")[1])
    print(f"{name}: {natoms} atoms -> {dst}")
    # the same relaxed system with other interaction settings
    for suffix, overrides in (variants or {}).items():
        with open(os.path.join(work, f"md_{suffix}.mdp"), "w") as fh:
            fh.write(mdp_with(MD + fep, overrides))
        gmx(["grompp", "-f", f"md_{suffix}.mdp", "-c", "em.gro", "-p", "topol.top", "-o", f"md_{suffix}.tpr", "-maxwarn", "10"],
            work)
        dst = os.path.join(OUT, f"{name}_{suffix}.tpr")
        with open(os.path.join(work, f"md_{suffix}.tpr"), "rb") as src, open(dst, "wb") as out:
            out.write(src.read())
        print(f"{name}_{suffix}: {natoms} atoms -> {dst}")


# BASELINE.json's configs[2] and configs[3] in kind (Gapsys soft-core with separate coul / vdw lambda
# paths; reaction-field with 40 lambda states and energy groups), on the 24.5 k-atom system, and an
# LJ-PME variant of the methane box
_LAM40 = " ".join(f"{v:.4f}" for v in np.linspace(0.0, 1.0, 40))
VARIANTS_C1 = {
    "ljpme": {"vdwtype": "PME", "lj-pme-comb-rule": "geometric", "ewald-rtol-lj": "1e-3"},
}
VARIANTS_C2 = {
    "gapsys": {"sc-function": "gapsys", "sc-gapsys-scale-linpoint-lj": "0.85", "sc-gapsys-scale-linpoint-q": "0.3",
               "sc-gapsys-sigma-lj": "0.3", "fep-lambdas": "",
               "coul-lambdas": "0.0 0.2 0.4 0.7 0.9 1.0 1.0 1.0", "vdw-lambdas": "0.0 0.0 0.1 0.3 0.5 0.7 0.9 1.0",
               "init-lambda-state": "3"},
    "rf": {"coulombtype": "reaction-field", "epsilon-rf": "78", "fep-lambdas": _LAM40, "init-lambda-state": "16",
           "sc-coul": "yes", "energygrps": "HEX SOL"},
}

if __name__ == "__main__":
    which = sys.argv[1:] or ["c1_methane", "c2_hexadecane"]
    if "c1_methane" in which:
        build("c1_methane", 3.0, methane, FEP_C1, VARIANTS_C1)
    if "c2_hexadecane" in which:
        build("c2_hexadecane", 6.3, hexadecane, FEP_C2, VARIANTS_C2)
    # the same solute in a 10 nm box (~100 k atoms, BASELINE configs[2] in size): for timings inside mdrun at a size where
    # the kernels, not the launches, dominate (tests/test_mdrun_nb_gpu_route.py)
    if "c3_hexadecane" in which:
        build("c3_hexadecane", 10.0, hexadecane, FEP_C2)
