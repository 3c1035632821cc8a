"""GPU parity of the drop-in path behind Nbnxm::gpu_* (b200nb_set_grid_atoms / _upload_pairlist / _copy_xq_grid / _get_f_grid):
the force kernel fed with the REFERENCE-built 8x8x8 pair list and grid-ordered atom data, exactly what gpu_init_pairlist /
gpu_init_atomdata receive (nbnxm_gpu_data_mgmt.cpp:251-311), against the reference's own outputs on that list
(nbnxn_kernel_gpu_ref) -- committed fixture -- and, where oracle/_ref travelled to this box, against the live reference."""
import os

import numpy as np
import pytest

import gmxapi_b200 as g
from gmxapi_b200 import lib as nb
from oracle import gmxref, oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_gpu_list_water_3k.npz")


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def upload(L, nbfp, box, rc, rlist, rlist_inner=0.0, **kw):
    h = nb.NbnxmGpu(0)
    h.set_params(nbfp, rc, rlist_outer=rlist, rlist_inner=rlist_inner, disp_cpot=-1.0 / rc ** 6, rep_cpot=-1.0 / rc ** 12,
                 epsfac=g.nblib.ONE_4PI_EPS0, **kw)
    h.set_box(box)
    h.set_grid_atoms(L["xq"], L["type"])
    h.upload_pairlist(0, L["sci"], L["cj4"], L["excl"])
    return h


def run(h, flags):
    h.clear_outputs()
    h.launch_force(0, flags)
    f = h.get_f_grid()
    fs, elj, eel = h.get_outputs()
    return f, fs, elj, eel


def test_reference_list_fixture(built):
    """3000-atom water, reaction field: list, atom data and expected outputs all come from the reference (fixture)."""
    G = np.load(GOLD)
    s = g.systems.named("water_3k")
    rc = float(G["rc"])
    h = upload(G, s.nbfp, G["box"], rc, rc, eeltype=nb.EEL_RF, k_rf=float(G["k_rf"]), c_rf=float(G["c_rf"]))
    f, fs, elj, eel = run(h, nb.FLAG_ENERGY | nb.FLAG_VIRIAL)
    real = G["atom_index"] >= 0
    assert np.all(f[~real] == 0)
    # per-slot forces against nbat->out[0].f of the reference's GPU-layout kernel on the same list
    assert relrms(f[real], G["f_grid"][real]) < 1e-5
    # in-range pair set: slots -> atoms through the reference's grid order
    p = h.pairs(rc)
    assert len(p) == int(G["npairs"])
    ai = G["atom_index"]
    pa = np.stack([ai[p[:, 0]], ai[p[:, 1]], p[:, 2]], 1)
    op = oracle.canonical_pairs(oracle.pair_set(s.x, s.box, rc, s.excl_off, s.excl_idx))
    assert np.array_equal(oracle.canonical_pairs(pa), op)
    m = np.ones(45, bool)
    m[nb.CENTRAL] = False
    assert np.abs(fs[m] - G["fshift"][m]).max() <= 1e-5 * np.abs(G["fshift"][m]).max()
    # the reference sums its energies in single precision (2e-4 bar as in test_gpu_parity)
    assert abs(elj - G["energies"][0]) <= 2e-4 * abs(G["energies"][0])
    assert abs(eel - G["energies"][1]) <= 2e-4 * abs(G["energies"][1])
    # new coordinates through gpu_copy_xq_to_gpu: a rigid translation leaves the forces unchanged
    xq = G["xq"].copy()
    xq[:, :3] += np.float32(0.25)
    h.copy_xq_grid(xq)
    f2, _, _, _ = run(h, 0)
    assert relrms(f2[real], f[real]) < 2e-5
    h.close()


@pytest.mark.skipif(not gmxref.available(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("name,rlist,rlist_inner", [("water_24k", 0.9, 0.0), ("water_3k", 1.0, 0.93)])
def test_reference_list_live_ewald(built, name, rlist, rlist_inner):
    """Ewald real space on a list the live reference builds here (buffered + dynamically pruned in the second case): forces in
    atom order against the plain-C oracle's analytical Ewald (the reference's GPU-layout CPU kernel only has the tabulated form)."""
    s = g.systems.named(name)
    rc = 0.9
    beta = float(np.float32(g.systems.ewald_beta(rc)))
    r = gmxref.RefNbnxm(s.x, s.box, s.types, s.q, s.nbfp, s.excl_off, s.excl_idx, rc=rc, rlist=rlist, eeltype=gmxref.EEL_EWALD_TAB,
                        ewaldcoeff=beta, kernel=gmxref.KERNEL_GPUREF, nthreads=1)
    L = r.gpu_list()
    ai = r.grid_order()
    h = upload(L, s.nbfp, s.box, rc, rlist, rlist_inner, eeltype=nb.EEL_EWALD, ewald_beta=beta)
    fg, fs, elj, eel = run(h, nb.FLAG_ENERGY | nb.FLAG_VIRIAL)
    real = ai >= 0
    f = np.zeros((len(s.types), 3), np.float32)
    f[ai[real]] = fg[real]
    fo, fso, evo, eco, npairs = oracle.forces(s.x, s.box, s.q, s.types, s.nbfp, rc, s.excl_off, s.excl_idx, eeltype=oracle.EEL_EWALD, beta=beta)
    assert h.pair_count(rc) == npairs
    assert relrms(f, fo) < 1e-5
    assert abs(elj - evo) <= 2e-5 * abs(evo) and abs(eel - eco) <= 2e-5 * abs(eco)
    if rlist_inner:
        # rolling prune of the uploaded list (gpu_launch_kernel_pruneonly): nothing inside rc may be lost
        for part in range(4):
            h.launch_prune(0, part, 4)
        assert h.pair_count(rc) == npairs
        f2 = run(h, 0)[0]
        assert relrms(f2[real], fg[real]) < 1e-6
    r.close()
    h.close()
