#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ (run in the dev container, where /root/reference
and oracle/_ref exist; the fixtures then travel to the GPU box, the reference tree does not).

 1. argon12_forces.json / spc_methanol_forces.json: the reference's own golden XMLs
    (api/nblib/tests/refdata/NBlibTest_{Argon,SpcMethanol}ForcesAreCorrect.xml) transcribed to JSON.
 2. ref_<system>_<eel>.npz: outputs of the UNMODIFIED reference code (oracle/_ref, built by
    oracle/build_ref.sh from the sources under /root/reference) on our seeded synthetic inputs:
      f (n,3) f32, fshift (45,3) f32, e_lj, e_el   -- CPU SIMD kernel (2xMM on AVX-512), the parity target
      npairs, pairs_sha256                           -- in-range non-excluded pair set at rc (canonical keys)
      grid_sha256, grid_dims                         -- GPU-geometry (8x8x8) grid atom order
 2b. ref_water_3k_triclinic_{ewald,rf}.npz: the same for the 3 k box sheared into a triclinic cell (`make_golden.py triclinic`).
 2c. ref_water_3k_fep_rf.npz: the reference's free-energy kernel on a perturbed pair list (`make_golden.py fep`).
 2d. ref_water_3k_fep_ljpme.npz: the same kernel with LJ-PME, both grid combination rules (`make_golden.py fep_ljpme`).
 2e. ref_water_3k_fep_twin.npz: the same kernel with rvdw < rcoulomb (`make_golden.py fep_twin`).
 3. ref_water_3k_vdw_<flavour>.npz: the reference's CPU SIMD kernels with an LJ force switch, an LJ potential switch
    and / or a VdW cut-off shorter than the Coulomb cut-off (Ewald electrostatics), once with the water charges and
    once with all charges zero (f_lj: Lennard-Jones forces alone, so that the modifier arithmetic is not hidden
    under the 100x larger Coulomb forces).
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/api/nblib/tests/refdata"
RC = 0.9


def xml_vectors(path):
    xml = open(path).read()
    return np.array([float(v) for v in re.findall(r'<Real Name="[XYZ]">([^<]+)</Real>', xml)]).reshape(-1, 3)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    for name, xml in (("argon12", "NBlibTest_ArgonForcesAreCorrect.xml"), ("spc_methanol", "NBlibTest_SpcMethanolForcesAreCorrect.xml")):
        f = xml_vectors(os.path.join(REF, xml))
        json.dump({"source": "api/nblib/tests/refdata/" + xml, "forces": f.tolist()},
                  open(os.path.join(HERE, name + "_forces.json"), "w"), indent=1)
    beta = float(np.float32(S.ewald_beta(RC)))
    k_rf, c_rf = S.rf_constants(RC)
    for sysname in ("water_3k", "water_24k"):
        s = S.named(sysname)
        for eel, kw in (("ewald", dict(eeltype=gmxref.EEL_EWALD_ANA, ewaldcoeff=beta)),
                        ("rf", dict(eeltype=gmxref.EEL_RF, k_rf=k_rf, c_rf=c_rf))):
            r = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, nthreads=4, **kw)
            f, fs, elj, eel_ = r.compute()
            keys = oracle.canonical_pairs(r.pair_set())
            r.close()
            g = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, nthreads=1,
                                kernel=gmxref.KERNEL_GPUREF, **kw)
            order = g.grid_order()
            dims = g.grid_dims()
            g.close()
            out = dict(fshift=fs, e_lj=np.float64(elj), e_el=np.float64(eel_), npairs=np.int64(len(keys)),
                       pairs_sha256=np.array(sha(keys)), grid_sha256=np.array(sha(order.astype(np.int32))),
                       grid_dims=np.array(dims[:2] + (dims[4],), np.int64), beta=np.float64(beta),
                       seed=np.int64(20261017))
            if sysname == "water_3k":
                out["f"] = f  # 36 kB; the 24k system keeps only checksums + a strided sample
            else:
                out["f_sample"] = f[::16].copy()
                out["f_sumsq"] = np.float64((f.astype(np.float64) ** 2).sum())
            np.savez_compressed(os.path.join(HERE, "ref_%s_%s.npz" % (sysname, eel)), **out)
            print(sysname, eel, len(keys), elj, eel_)


VDW_FLAVOURS = {  # name: (vdw_modifier, rvdw, rvdw_switch)
    "twin": (0, 0.8, 0.0),
    "fswitch": (1, 0.9, 0.75),
    "pswitch": (2, 0.9, 0.75),
    "fswitch_twin": (1, 0.8, 0.65),
}


def vdw_flavours():
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    beta = float(np.float32(S.ewald_beta(RC)))
    s = S.named("water_3k")
    for name, (mod, rvdw, rsw) in VDW_FLAVOURS.items():
        k = oracle.vdw_modifier_constants(mod, rvdw, rsw)
        out = dict(beta=np.float64(beta), vdw_modifier=np.int64(mod), rvdw=np.float64(rvdw), rvdw_switch=np.float64(rsw))
        for tag, q in (("", s.q), ("_lj", np.zeros_like(s.q))):
            r = gmxref.RefNbnxm(s.x, s.box, s.types, q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, nthreads=4,
                                eeltype=gmxref.EEL_EWALD_ANA, ewaldcoeff=beta, rvdw=rvdw if rvdw < RC else 0.0,
                                vdw_modifier=mod, rvdw_switch=rsw, modifier_constants=k)
            f, fs, elj, eel_ = r.compute()
            r.close()
            out["f" + tag], out["fshift" + tag] = f, fs
            out["e_lj" + tag], out["e_el" + tag] = np.float64(elj), np.float64(eel_)
        np.savez_compressed(os.path.join(HERE, "ref_water_3k_vdw_%s.npz" % name), **out)
        print(name, out["e_lj"], out["e_el"], out["e_lj_lj"])


def ljpme_flavours():
    """ref_water_3k_ljpme_{geom,lb}.npz: LJ-PME real-space kernels of the reference (SIMD for the geometric grid rule, plain C
    for Lorentz-Berthelot: the only kernel that has it) on the 3 k water box with LJ on the hydrogens too
    (systems.nbfp_two_lj_types), with the water charges and with all charges zero."""
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    beta = float(np.float32(S.ewald_beta(RC)))
    bl = float(np.float32(S.ewald_beta_lj(RC)))
    sh = oracle.lj_ewald_shift(bl, RC)
    s = S.named("water_3k")
    nbfp = S.nbfp_two_lj_types()
    for name, ljpme, kern in (("geom", 1, None), ("lb", 2, gmxref.KERNEL_PLAINC)):
        out = dict(beta=np.float64(beta), ewaldcoeff_lj=np.float64(bl), sh_lj_ewald=np.float64(sh), ljpme=np.int64(ljpme), nbfp=nbfp)
        for tag, q in (("", s.q), ("_lj", np.zeros_like(s.q))):
            r = gmxref.RefNbnxm(s.x, s.box, s.types, q, nbfp, s.excl_off, s.excl_idx, rc=RC, nthreads=4, kernel=kern,
                                eeltype=gmxref.EEL_EWALD_ANA, ewaldcoeff=beta, comb_rule=ljpme, ljpme=ljpme,
                                ewaldcoeff_lj=bl, sh_lj_ewald=sh)
            f, fs, elj, eel_ = r.compute()
            r.close()
            out["f" + tag], out["fshift" + tag] = f, fs
            out["e_lj" + tag], out["e_el" + tag] = np.float64(elj), np.float64(eel_)
        np.savez_compressed(os.path.join(HERE, "ref_water_3k_ljpme_%s.npz" % name), **out)
        print(name, out["e_lj"], out["e_el"], out["e_lj_lj"])


def triclinic():
    """ref_water_3k_triclinic_{ewald,rf}.npz: the reference on the 3 k water box sheared into a triclinic cell
    (gmxapi_b200.systems.sheared: box[YY][XX] = 0.25 L, box[ZZ][XX] = -0.2 L, box[ZZ][YY] = 0.3 L)."""
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    beta = float(np.float32(S.ewald_beta(RC)))
    k_rf, c_rf = S.rf_constants(RC)
    s = S.sheared(S.named("water_3k"))
    for eel, kw in (("ewald", dict(eeltype=gmxref.EEL_EWALD_ANA, ewaldcoeff=beta)), ("rf", dict(eeltype=gmxref.EEL_RF, k_rf=k_rf, c_rf=c_rf))):
        r = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=RC, nthreads=4, box_offdiag=s.box_offdiag, **kw)
        f, fs, elj, eel_ = r.compute()
        keys = oracle.canonical_pairs(r.pair_set())
        r.close()
        out = dict(f=f, fshift=fs, e_lj=np.float64(elj), e_el=np.float64(eel_), npairs=np.int64(len(keys)), pairs_sha256=np.array(sha(keys)),
                   beta=np.float64(beta), box=s.box, box_offdiag=s.box_offdiag, x_sha256=np.array(sha(s.x)), seed=np.int64(20261017))
        np.savez_compressed(os.path.join(HERE, "ref_water_3k_triclinic_%s.npz" % eel), **out)
        print("triclinic", eel, len(keys), elj, eel_)


def fep():
    """ref_water_3k_fep_rf.npz: the reference's free-energy kernel (gmxlib/nonbonded/nb_free_energy.cpp, compiled into oracle/_ref) on the
    perturbed pair list of gmxapi_b200.systems.perturbed_water, reaction field, for the soft-core settings of systems.FEP_CASES."""
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    s, pert, tA, tB, qA, qB, tm, qm = S.perturbed_water()
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    sv = oracle.shift_vectors(s.box)
    k_rf, c_rf = S.rf_constants(RC, eps_rf=1.0)
    out = dict(list_sha256=np.array(sha(np.concatenate([a.astype(np.int64).ravel() for a in lst]))), npairs=np.int64(len(lst[3])), nri=np.int64(len(lst[0])))
    # the reference's OWN perturbed pair lists (nbnxm/pairlist.cpp make_fep_list on the GPU-layout list, through the search of
    # oracle/_ref with the perturbed atoms flagged), restricted to the pairs within rlist (the reference keeps a buffer beyond it):
    # hash and count of the canonical pair keys, for rlist = rc and rlist > rc
    for rl in (0.9, 1.0):
        r = gmxref.RefNbnxm(s.x, s.box, tm, qm, s.nbfp, s.excl_off, s.excl_idx, RC, rlist=rl, eeltype=gmxref.EEL_RF, k_rf=k_rf, c_rf=c_rf,
                            kernel=gmxref.KERNEL_GPUREF, perturbed=pert.astype(np.uint8), nthreads=1)
        ref = gmxref.fep_list(r)
        keys = oracle.fep_list_canonical(ref, oracle.fep_list_within(ref, s.x, s.box, rl))
        out["ref_list_sha256_%.1f" % rl], out["ref_list_npairs_%.1f" % rl], out["ref_list_npairs_all_%.1f" % rl] = np.array(sha(keys)), np.int64(len(keys)), np.int64(len(ref[3]))
        print("reference fep list rlist", rl, len(ref[3]), "pairs,", len(keys), "within rlist")
    for name, kw in S.FEP_CASES.items():
        f, fs, o4 = gmxref.fep_kernel(s.x, sv, s.nbfp, tA, tB, qA, qB, *lst, RC, k_rf=k_rf, c_rf=c_rf, **kw)
        out["f_" + name], out["fshift_" + name], out["out4_" + name] = f, fs, np.array(o4, np.float64)
        print("fep", name, o4)
    # LJ potential switch 0.75 -> 0.9 (eintmodPOTSWITCH: on the soft-cored distance, nb_free_energy.cpp:613-625)
    for name, kw in S.FEP_CASES.items():
        f, fs, o4 = gmxref.fep_kernel(s.x, sv, s.nbfp, tA, tB, qA, qB, *lst, RC, k_rf=k_rf, c_rf=c_rf, rvdw_switch=0.75, **kw)
        out["f_pswitch_" + name], out["fshift_pswitch_" + name], out["out4_pswitch_" + name] = f, fs, np.array(o4, np.float64)
        print("fep pswitch", name, o4)
    np.savez_compressed(os.path.join(HERE, "ref_water_3k_fep_rf.npz"), **out)
    # the same with Ewald electrostatics: ref_water_3k_fep_ewald.npz
    import math
    beta = float(np.float32(S.ewald_beta(RC)))
    sh = float(np.float32(math.erfc(beta * RC) / RC))
    oute = dict(list_sha256=out["list_sha256"], beta=np.float64(beta), sh_ewald=np.float64(sh))
    for name, kw in S.FEP_CASES.items():
        f, fs, o4 = gmxref.fep_kernel(s.x, sv, s.nbfp, tA, tB, qA, qB, *lst, RC, ewaldcoeff=beta, sh_ewald=sh, **kw)
        oute["f_" + name], oute["fshift_" + name], oute["out4_" + name] = f, fs, np.array(o4, np.float64)
        print("fep ewald", name, o4)
    np.savez_compressed(os.path.join(HERE, "ref_water_3k_fep_ewald.npz"), **oute)


def fep_ljpme():
    """ref_water_3k_fep_ljpme.npz: the reference's free-energy kernel with vdwtype = PME (the grid part of the dispersion subtracted
    from its cubic-spline table, nb_free_energy.cpp:725-770; fr->ljpme_c6grid restated in oracle/ref_harness.cpp from the static
    make_ljpme_c6grid) and Ewald electrostatics, both grid combination rules, on systems.perturbed_water_ljpme."""
    import math
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    s, pert, tA, tB, qA, qB, tm, qm, nbfp = S.perturbed_water_ljpme()
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    sv = oracle.shift_vectors(s.box)
    beta = float(np.float32(S.ewald_beta(RC)))
    sh = float(np.float32(math.erfc(beta * RC) / RC))
    bl = float(np.float32(S.ewald_beta_lj(RC)))
    shlj = oracle.lj_ewald_shift(bl, RC)
    out = dict(list_sha256=np.array(sha(np.concatenate([a.astype(np.int64).ravel() for a in lst]))), beta=np.float64(beta), sh_ewald=np.float64(sh),
               ewaldcoeff_lj=np.float64(bl), sh_lj_ewald=np.float64(shlj), nbfp=nbfp)
    for rule, tag in ((1, "geom"), (2, "lb")):
        for name, kw in S.FEP_CASES.items():
            f, fs, o4 = gmxref.fep_kernel(s.x, sv, nbfp, tA, tB, qA, qB, *lst, RC, ewaldcoeff=beta, sh_ewald=sh, ljpme=rule, ewaldcoeff_lj=bl,
                                          sh_lj_ewald=shlj, **kw)
            out["f_%s_%s" % (tag, name)], out["fshift_%s_%s" % (tag, name)], out["out4_%s_%s" % (tag, name)] = f, fs, np.array(o4, np.float64)
            print("fep ljpme", tag, name, o4)
    # the same pairs with cut-off LJ, for the size of the grid part
    f, fs, o4 = gmxref.fep_kernel(s.x, sv, nbfp, tA, tB, qA, qB, *lst, RC, ewaldcoeff=beta, sh_ewald=sh, **S.FEP_CASES["sc1"])
    out["f_cut_sc1"], out["out4_cut_sc1"] = f, np.array(o4, np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_water_3k_fep_ljpme.npz"), **out)


def fep_twin():
    """ref_water_3k_fep_twin.npz: the reference's free-energy kernel with rvdw = 0.8 < rcoulomb = 0.9 (what PME load balancing leaves
    behind: nb_free_energy.cpp:300-301, :564-587 cut the two interactions separately) on systems.perturbed_water_ljpme."""
    import math
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    s, pert, tA, tB, qA, qB, tm, qm, nbfp = S.perturbed_water_ljpme()
    lst = oracle.fep_pair_list(s.x, s.box, RC, pert, s.excl_off, s.excl_idx)
    sv = oracle.shift_vectors(s.box)
    rvdw = 0.8
    beta = float(np.float32(S.ewald_beta(RC)))
    sh = float(np.float32(math.erfc(beta * RC) / RC))
    bl = float(np.float32(S.ewald_beta_lj(rvdw)))
    shlj = oracle.lj_ewald_shift(bl, rvdw)
    out = dict(list_sha256=np.array(sha(np.concatenate([a.astype(np.int64).ravel() for a in lst]))), beta=np.float64(beta), sh_ewald=np.float64(sh),
               ewaldcoeff_lj=np.float64(bl), sh_lj_ewald=np.float64(shlj), rvdw=np.float64(rvdw))
    for tag, rule, sw, case in S.FEP_TWIN:
        kw = dict(S.FEP_CASES[case], ewaldcoeff=beta, sh_ewald=sh, rvdw=rvdw, rvdw_switch=sw)
        if rule:
            kw.update(ljpme=rule, ewaldcoeff_lj=bl, sh_lj_ewald=shlj)
        f, fs, o4 = gmxref.fep_kernel(s.x, sv, nbfp, tA, tB, qA, qB, *lst, RC, **kw)
        out["f_" + tag], out["fshift_" + tag], out["out4_" + tag] = f, fs, np.array(o4, np.float64)
        print("fep twin", tag, o4)
    np.savez_compressed(os.path.join(HERE, "ref_water_3k_fep_twin.npz"), **out)


BONDED_TRICLINIC = ((3.1, 0.0, 0.0), (0.6, 2.9, 0.0), (-0.5, 0.7, 3.3))


def bonded():
    """ref_bonded_chains.npz: the reference's CPU functions (listed_forces/bonded.cpp calculateSimpleBond, pairs.cpp do_pairs,
    compiled into oracle/_ref) for the interaction types its GPU bonded module covers, on gmxapi_b200.systems.bonded_chains in the
    rectangular box and wrapped into a triclinic one."""
    import gmxapi_b200.systems as S
    from oracle import gmxref, oracle
    out = dict(box_triclinic=np.array(BONDED_TRICLINIC, np.float32))
    for tag, bm in (("rect", None), ("tric", BONDED_TRICLINIC)):
        s = S.bonded_chains(box_matrix=bm)
        B, x = (np.diag(s["box"]) if bm is None else np.array(bm)).astype(np.float32), s["x"]
        out["x_sha256_" + tag] = np.array(sha(x))
        for kind in oracle.BONDED_KINDS:
            d = s[kind]
            f, fs, e = gmxref.bonded(kind, d["iatoms"], d["params"], x, s["q"], B)
            out["f_%s_%s" % (tag, kind)], out["fshift_%s_%s" % (tag, kind)], out["e_%s_%s" % (tag, kind)] = f, fs, np.float64(e)
            print("bonded", tag, kind, len(d["iatoms"]), e)
    np.savez_compressed(os.path.join(HERE, "ref_bonded_chains.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bonded":
        bonded()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ljpme":
        ljpme_flavours()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vdw":
        vdw_flavours()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "triclinic":
        triclinic()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fep":
        fep()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fep_ljpme":
        fep_ljpme()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fep_twin":
        fep_twin()
        sys.exit(0)
    main()
