"""The drop-in itself: the reference's own C++ driver (XML reader, set-up, archiver) with tasks 1-9,11
replaced by GpuTask objects that call libmpmgpu (nairn_mpm_fea_b200/host/GpuTasks.cpp), run on the same XML
input as the unmodified reference CLI; every binary archive both write must agree (element ids exactly,
doubles to 1e-7 of the column max).  Needs the prebuilt oracle/_ref/NairnMPM and host/_build/NairnMPM_gpu."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests import inputs
from tests.archive import list_archives, read_archive

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "NairnMPM")
GPU = os.path.join(ROOT, "nairn_mpm_fea_b200", "host", "_build", "NairnMPM_gpu")

CASES = {
    "block3d": (inputs.block3d(ncell=6, margin=3, maxtime=0.03).replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>",
                                                                          "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), 6 ** 3 * 8, "res/blk."),
    "disks2d": (inputs.disks2d(analysis=10, maxtime=2.0, archive_ms=0.5), None, "res/disks."),
    # config 3 family: Neo-Hookean, lCPDI, XPIC(2) scheduled by the reference's own PeriodicXPIC custom task
    "block3d_neo_lcpdi_xpic2": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, gimp="lCPDI", material=inputs.neohookean_material(), vz=-8.0e3, vx=2.0e3,
                                               custom_tasks=inputs.periodic_xpic(2, False, 1))
                                .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), None, "res/blk."),
    # grid velocity BCs that vary in time: a function of time and position on the bottom plane, a linear ramp that starts
    # late, and a skewed (xy) condition on one side -- the adapter re-evaluates the list every step
    "block3d_bc_functions": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=-3.0e3)
                             .replace('<DisBC dir="3" vel="0"/>', '<DisBC dir="3" style="6" function="400*sin(200*t)*(1+0.1*x)"/>')
                             .replace("</GridBCs>", '<BCBox xmin="-1" xmax="3.01" ymin="-1" ymax="20" zmin="-1" zmax="20">'
                                                    '<DisBC dir="12" angle="30" style="2" vel="50000" time="0.01"/></BCBox></GridBCs>')
                             .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), None, "res/blk."),
    # a rigid piston whose velocity follows setting functions of time and position (evaluated by the host every step)
    "block3d_rigid_piston_functions": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0,
                                                      rigid=("piston", 5, (0.0, 0.0, 0.0), ("300*sin(40*t)*(1+0.05*y)", "-9000*(1-exp(-t/0.004))")))
                                       .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), None, "res/blk."),
    # particle load BCs (MatPtLoadBC) re-evaluated every step: a constant load on the top layer, a linear ramp that starts late on one
    # side and a sine load on another face; two BCs overlap on the corner particles
    "block3d_particle_loads": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0)
                               .replace("</GridBCs>", "</GridBCs><ParticleBCs>"
                                        '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="6.5" zmax="20"><LoadBC dir="3" style="1" load="-0.5"/></BCBox>'
                                        '<BCBox xmin="6.5" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="20"><LoadBC dir="1" style="2" load="40" time="0.005"/></BCBox>'
                                        '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="3.5" zmin="-1" zmax="20"><LoadBC dir="2" style="3" load="0.3" time="300"/></BCBox>'
                                        '<BCBox xmin="-1" xmax="3.5" ymin="-1" ymax="20" zmin="-1" zmax="20"><LoadBC dir="1" style="6" function="0.4*sin(300*t)"/></BCBox>'
                                        "</ParticleBCs>")
                               .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), None, "res/blk."),
    # multimaterial mode: the two disks keep their own velocity fields and meet with Coulomb friction
    "disks2d_multimaterial_friction": (inputs.oblique_disks(inputs.disks2d(analysis=10, vel=4000.0, vmax=11.0, gap=0.0, maxtime=0.6, archive_ms=0.15,
                                                            extra_header=inputs.multimaterial(2, 0.3))), None, "res/disks."),
    "blocks3d_multimaterial_position": (inputs.blocks3d_contact(inputs.multimaterial(0, 0.25, 0.8), materials=2)
                                        .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                        .replace('max="1.0"', 'max="0.03"'), None, "res/blk."),
    # a plate of rigid contact particles (RigidMaterial, SetDirection 8) whose velocity follows a setting function pushes a disk
    "disks2d_rigid_contact_plate": (inputs.rigid_contact_plate(inputs.disks2d(analysis=10, vel=3000.0, vmax=11.0, gap=0.0, maxtime=0.6, archive_ms=0.15,
                                                               extra_header=inputs.multimaterial(0, 0.3, None, ' RigidBias="10"')))
                                    .replace("<SetDirection>8</SetDirection>", "<SetDirection>8</SetDirection><SettingFunction>-2000*(1+0.3*sin(20*t))</SettingFunction>"
                                                                               "<SettingFunction2>700</SettingFunction2>"), None, "res/disks."),
    # SCGL hardening (pressure- and temperature-dependent shear modulus) on a block that starts 80 K above the stress-free temperature
    "block3d_scgl_thermal_offset": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, material=inputs.isoplastic_hardening_material("SCGL", yieldMax=400.0, GPpG0=4.0e-4, GTpG0=-1.0e-3),
                                                   vz=-4.0e4, vx=5.0e3, extra_header="<StressFreeTemp>300</StressFreeTemp>").replace('<Body ', '<Body temp="380" ', 1)
                                    .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                    .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYYNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # conduction with a hot rigid piston (<SetTemperature/>: the nodes it touches are held at its temperature) over a block whose bottom
    # plane carries a grid temperature BC
    "block3d_conduction_rigid_hot_piston": (inputs.conduction(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0, rigid=("piston", 4, (0.0, 0.0, -3.0e3)))
                                                              .replace("<SetDirection>4</SetDirection>", "<SetDirection>4</SetDirection><SetTemperature/>"),
                                                              (300.0, 600.0), (4000.0,), (700.0,))
                                            .replace("</GridBCs>", '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="3.01"><TempBC value="350"/></BCBox></GridBCs>')
                                            .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                            .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # particle traction BCs: a pressure that ramps up on the top face (normal to the deformed face) and a constant shear on one side
    "block3d_traction_pressure_ramp": (inputs.particle_bcs(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0), [
        ('<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="6.5" zmax="20">', '<TractionBC dir="11" face="6" style="2" stress="-400"/>'),
        ('<BCBox xmin="6.5" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="20">', '<TractionBC dir="3" face="2" style="1" stress="2"/>'),
        ('<BCBox xmin="-1" xmax="20" ymin="-1" ymax="3.5" zmin="-1" zmax="20">', '<TractionBC dir="2" face="1" style="6" function="3*sin(200*t)"/>')])
                                       .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"), None, "res/blk."),
    "disks2d_traction_lcpdi": (inputs.particle_bcs(inputs.oblique_disks(inputs.disks2d(analysis=10, gimp="lCPDI", vel=1000.0, vmax=11.0, gap=0.0, maxtime=0.6, archive_ms=0.15)), [
        ('<BCLine x1="-12" y1="-11" x2="-12" y2="11" tolerance="3">', '<TractionBC dir="11" face="4" style="1" stress="-0.005"/>'),
        ('<BCLine x1="-20" y1="6" x2="20" y2="6" tolerance="2">', '<TractionBC dir="12" face="3" style="2" stress="0.01"/>')]), None, "res/disks."),
    # conduction with a heat flux that ramps up on the top face of the block and a constant one leaving through a side
    "block3d_conduction_heat_flux": (inputs.particle_bcs(inputs.conduction(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=-2.0e3, vx=1.0e3), (300.0,), (4000.0,), (700.0,)), [
        ('<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="6.5" zmax="20">', '<HeatFluxBC dir="1" face="6" style="2" value="4e9"/>'),
        ('<BCBox xmin="6.5" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="20">', '<HeatFluxBC dir="1" face="2" style="1" value="-2e7"/>'),
        ('<BCBox xmin="-1" xmax="20" ymin="-1" ymax="3.5" zmin="-1" zmax="20">', '<HeatFluxBC dir="1" face="1" style="6" function="1e7*(1-cos(300*t))"/>')])
                                     .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                     .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # conduction under a mechanical FMPM(2) update scheduled by PeriodicXPIC
    "block3d_conduction_fmpm2": (inputs.conduction(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=-2.0e3, vx=1.0e3, custom_tasks=inputs.periodic_xpic(2, True, 1))
                                                   .replace("<alpha>0</alpha>", "<alpha>50</alpha>"), (300.0,), (4000.0,), (700.0,))
                                 .replace("</GridBCs>", '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="3.01"><TempBC value="450"/></BCBox></GridBCs>')
                                 .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                 .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # ... and with the piston's temperature set by a value function of time and position (evaluated by the host every step)
    "block3d_conduction_rigid_piston_value_function": (inputs.conduction(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0, rigid=("piston", 4, (0.0, 0.0, -3.0e3)))
                                                                         .replace("<SetDirection>4</SetDirection>", "<SetDirection>4</SetDirection><SetTemperature/><ValueFunction>500+6000*t+10*x</ValueFunction>"),
                                                                         (300.0, 600.0), (4000.0,), (700.0,))
                                                       .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                                       .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # adiabatic coupling: a Johnson-Cook block heats itself by plastic work (thermal softening), no transport task
    "block3d_adiabatic_johnsoncook": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, material=inputs.isoplastic_hardening_material("JohnsonCook", Djc=0.01), vz=-4.0e4,
                                                     extra_header="<StressFreeTemp>300</StressFreeTemp>").replace("</JANFEAInput>", "<Thermal><EnergyCoupling/></Thermal></JANFEAInput>")
                                      .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                      .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYYNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # a block that starts 60 K above its stress-free temperature (thermal expansion, no transport task) and yields
    "block3d_thermal_offset_isoplastic": (inputs.block3d(ncell=4, margin=3, maxtime=0.03, material=inputs.isoplastic_material(yld=5.0), vz=-2.0e4,
                                                         extra_header="<StressFreeTemp>300</StressFreeTemp>").replace('<Body ', '<Body temp="360" ', 1)
                                          .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                          .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYYNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # conduction with nodal temperature BCs: a constant one on the bottom plane and a ramp that starts late on one side
    "block3d_conduction_temperature_bcs": (inputs.conduction(inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=-2.0e3, vx=1.0e3).replace("<alpha>0</alpha>", "<alpha>50</alpha>"),
                                                             (300.0,), (4000.0,), (700.0,))
                                           .replace("</GridBCs>", '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="4.01"><TempBC value="450"/></BCBox>'
                                                                  '<BCBox xmin="-1" xmax="4.01" ymin="-1" ymax="20" zmin="-1" zmax="20"><TempBC style="2" value="20000" time="0.005"/></BCBox></GridBCs>')
                                           .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
                                           .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/blk."),
    # heat conduction between two disks of different temperature, conductivity and heat capacity; the archives carry the
    # particle temperature
    "disks2d_conduction": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=2000.0, vmax=11.0, gap=0.0, alpha=60.0, maxtime=0.6, archive_ms=0.15)),
                                             (380.0, 290.0), (2000.0, 500.0), (800.0, 1500.0))
                           .replace("<MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>", "<MPMArchiveOrder>iYYYYNNYNNNNYNNNNY</MPMArchiveOrder>"), None, "res/disks."),
    # config 4 family: IsoPlasticity bar on a plate of rigid-BC particles
    "block3d_isoplastic_rigid_wall": (inputs.block3d(ncell=4, margin=3, maxtime=0.02, material=inputs.isoplastic_material(), vz=-4.0e4, bc=False,
                                                     rigid=("wall", 4, (0.0, 0.0, 0.0)))
                                      .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.005</ArchiveTime>"), None, "res/blk."),
}


def run(binary, xml, extra=()):
    d = tempfile.mkdtemp(prefix="dropin_")
    path = os.path.join(d, "in.fmcmd")
    open(path, "w").write(xml)
    p = subprocess.run([binary, *extra, path], cwd=d, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    return d, p.stdout


@pytest.mark.parametrize("mode", ["tasks", "fused"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_driver_with_gpu_tasks_matches_reference(case, mode):
    """mode tasks: every reference task replaced by the entry point of the same name (per-task kernels);
    mode fused: `NairnMPM_gpu -fused`, the whole step in one call (fused kernels for 3D uGIMP, else per-task)."""
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/NairnMPM or host/_build/NairnMPM_gpu not built")
    xml, npart, root = CASES[case]
    dref, out_ref = run(REF, xml, ("-np", "4"))
    dgpu, out_gpu = run(GPU, xml, ("-fused",) if mode == "fused" else ())
    assert "GPU TASKS" in out_gpu and ("whole-step" in out_gpu) == (mode == "fused")
    assert "GPU ARCHIVES" in out_gpu, "the archives of this run should have been packed on the device (SURVEY.md 8(f) row 1)"
    if npart is None:
        for ln in out_ref.splitlines():
            if "Number of Material Points:" in ln:
                npart = int(ln.split(":")[1].split()[0])
                break
        assert npart, "could not find particle count in the reference report"
    a_ref = list_archives(os.path.join(dref, root))
    a_gpu = list_archives(os.path.join(dgpu, root))
    assert [s for s, _ in a_ref] == [s for s, _ in a_gpu] and len(a_ref) >= 3, (a_ref, a_gpu)
    for (step, fr), (_, fg) in zip(a_ref, a_gpu):
        r, g = read_archive(fr, npart), read_archive(fg, npart)
        assert np.array_equal(r["elem"], g["elem"]), "element ids differ at step %d" % step
        assert np.array_equal(r["tail"], g["tail"]) and np.array_equal(r["mat"], g["mat"])
        scale = np.maximum(np.max(np.abs(r["doubles"]), axis=0), 1e-300)
        err = np.max(np.abs(r["doubles"] - g["doubles"]) / scale)
        assert err < 1e-7, "archive at step %d differs: %.3e" % (step, err)


@pytest.mark.parametrize("case,ngpus", [("block3d_two_slabs", 2), ("block3d_rigid_wall_two_slabs", 2)])
def test_drop_in_on_several_gpus_matches_reference(case, ngpus):
    """`NairnMPM_gpu -gpus N`: the reference driver with one slab context per GPU (cell planes along z, NCCL halo and migrant
    exchanges inside libmpmgpu, one host thread per GPU); every archive against the unmodified reference CLI.  The block is
    thrown downwards and sideways so that particles change slabs during the run."""
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip("needs %d GPUs" % ngpus)
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/NairnMPM or host/_build/NairnMPM_gpu not built")
    if case == "block3d_two_slabs":
        xml = inputs.block3d(ncell=8, margin=3, maxtime=0.06, E=100.0, vz=-2.5e4, vx=4.0e3)
    else:
        xml = inputs.block3d(ncell=6, margin=3, maxtime=0.04, material=inputs.isoplastic_material(), vz=-4.0e4, bc=False, rigid=("wall", 4, (0.0, 0.0, 0.0)))
    xml = xml.replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>")
    dref, out_ref = run(REF, xml, ("-np", "4"))
    dgpu, out_gpu = run(GPU, xml, ("-gpus", str(ngpus)))
    assert "GPU SLABS: %d GPUs" % ngpus in out_gpu, out_gpu[-1500:]
    moved = [int(ln.split()[2]) for ln in out_gpu.splitlines() if ln.startswith("GPU SLABS:") and "changed slabs" in ln]
    assert moved and moved[0] > 0, "no particle changed slabs: the test does not exercise the migration"
    npart = None
    for ln in out_ref.splitlines():
        if "Number of Material Points:" in ln:
            npart = int(ln.split(":")[1].split()[0])
            break
    a_ref = list_archives(os.path.join(dref, "res/blk."))
    a_gpu = list_archives(os.path.join(dgpu, "res/blk."))
    assert [s for s, _ in a_ref] == [s for s, _ in a_gpu] and len(a_ref) >= 3, (a_ref, a_gpu)
    for (step, fr), (_, fg) in zip(a_ref, a_gpu):
        r, g = read_archive(fr, npart), read_archive(fg, npart)
        assert np.array_equal(r["elem"], g["elem"]), "element ids differ at step %d" % step
        assert np.array_equal(r["tail"], g["tail"]) and np.array_equal(r["mat"], g["mat"])
        scale = np.maximum(np.max(np.abs(r["doubles"]), axis=0), 1e-300)
        err = np.max(np.abs(r["doubles"] - g["doubles"]) / scale)
        assert err < 1e-7, "archive at step %d differs: %.3e" % (step, err)


def test_device_packed_archives_equal_the_host_writers():
    """Same input twice: records packed on the device (default) and `-hostoutput` (full download, the reference's own
    ArchiveResults on mpm[]).  Two runs differ in the last bits (FP64 atomics add in another order from run to run), so the
    files are compared field by field to 1e-9; that the packer writes the same BYTES as the host writer from one state is
    tests/test_zzz_archive_gpu.py.  The global-quantity files are compared to their printed precision."""
    if not os.path.exists(GPU):
        pytest.skip("host/_build/NairnMPM_gpu not built")
    xml = (inputs.block3d(ncell=6, margin=3, maxtime=0.03, material=inputs.isoplastic_material(), vz=-4.0e4)
           .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.004</ArchiveTime>"
                    "<GlobalArchiveTime units=\"ms\">0.002</GlobalArchiveTime><GlobalArchive type=\"Kinetic Energy\"/>"
                    "<GlobalArchive type=\"Strain Energy\"/><GlobalArchive type=\"Plastic Energy\"/><GlobalArchive type=\"szz\"/>"
                    "<GlobalArchive type=\"velz\"/><GlobalArchive type=\"Fzz\"/><GlobalArchive type=\"Step number\"/>"))
    npart = 6 ** 3 * 8
    ddev, out_dev = run(GPU, xml, ("-fused",))
    dhost, out_host = run(GPU, xml, ("-fused", "-hostoutput"))
    assert "GPU ARCHIVES" in out_dev and "GPU ARCHIVES" not in out_host
    a_dev, a_host = list_archives(os.path.join(ddev, "res/blk.")), list_archives(os.path.join(dhost, "res/blk."))
    assert [s for s, _ in a_dev] == [s for s, _ in a_host] and len(a_dev) >= 5
    for (step, fd), (_, fh) in zip(a_dev, a_host):
        assert open(fd, "rb").read(64) == open(fh, "rb").read(64), "header of archive %d differs" % step
        d, h = read_archive(fd, npart), read_archive(fh, npart)
        assert np.array_equal(d["elem"], h["elem"]) and np.array_equal(d["tail"], h["tail"]) and np.array_equal(d["mat"], h["mat"])
        scale = np.maximum(np.max(np.abs(h["doubles"]), axis=0), 1e-300)
        assert np.max(np.abs(d["doubles"] - h["doubles"]) / scale) < 1e-9, "archive %d differs" % step
    gd, gh = open(os.path.join(ddev, "res/blk.global")).read(), open(os.path.join(dhost, "res/blk.global")).read()
    assert gd.count("\n") >= 8
    rows_d = [ln.split("\t") for ln in gd.splitlines() if not ln.startswith("#")]
    rows_h = [ln.split("\t") for ln in gh.splitlines() if not ln.startswith("#")]
    assert len(rows_d) == len(rows_h)
    for rd, rh in zip(rows_d, rows_h):
        assert len(rd) == len(rh)
        for x, y in zip(rd, rh):
            assert abs(float(x) - float(y)) <= 2e-6 * max(abs(float(y)), 1e-300) + 1e-30, (rd, rh)


@pytest.mark.parametrize("mode", ["tasks", "fused", "fused_hostoutput"])
def test_reaction_force_global_quantities_match_reference(mode):
    """"reactionx/y/z" global quantities (GlobalQuantity.cpp:971-986) by BC id: the bottom plane's grid BCs (id -1), the BCs a
    rigid piston makes (id = its material number, 2) and all of them (0).  The device keeps NodalVelBC::freaction per BC
    (mpmgpu_track_reactions); the adapter writes them into the host's BC objects and the reference's own code sums them."""
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/NairnMPM or host/_build/NairnMPM_gpu not built")
    xml = (inputs.block3d(ncell=4, margin=3, maxtime=0.03, E=100.0, vz=0.0, rigid=("piston", 7, (1.0e3, -5.0e2, -6.0e3)))
           .replace('<DisBC dir="3" vel="0"/>', '<DisBC dir="3" vel="0" id="-1"/>')
           .replace("<ArchiveTime units=\"ms\">1000</ArchiveTime>", "<ArchiveTime units=\"ms\">0.01</ArchiveTime>"
                    "<GlobalArchiveTime units=\"ms\">0.002</GlobalArchiveTime><GlobalArchive type=\"reactionz\" material=\"-1\"/>"
                    "<GlobalArchive type=\"reactionz\" material=\"2\"/><GlobalArchive type=\"reactionx\" material=\"2\"/>"
                    "<GlobalArchive type=\"reactiony\"/><GlobalArchive type=\"reactionz\"/><GlobalArchive type=\"Kinetic Energy\"/>"
                    "<GlobalArchive type=\"Grid Kinetic Energy\"/>"))        # (0.5 |pk|^2 / m over the host's nodes, which the adapter fills from the device)
    assert 'id="-1"' in xml
    dref, out_ref = run(REF, xml, ("-np", "4"))
    extra = {"tasks": (), "fused": ("-fused",), "fused_hostoutput": ("-fused", "-hostoutput")}[mode]
    dgpu, out_gpu = run(GPU, xml, extra)
    assert "GPU TASKS" in out_gpu
    rows_r = [ln.split("\t") for ln in open(os.path.join(dref, "res/blk.global")).read().splitlines() if not ln.startswith("#")]
    rows_g = [ln.split("\t") for ln in open(os.path.join(dgpu, "res/blk.global")).read().splitlines() if not ln.startswith("#")]
    assert len(rows_r) == len(rows_g) and len(rows_r) >= 10
    cols = np.array([[float(x) for x in r] for r in rows_r])
    assert np.all(np.abs(cols[:, 1:8]).max(axis=0) > 0), "a reaction or energy column is zero throughout: the input does not exercise it"
    for rr, rg in zip(rows_r, rows_g):
        assert len(rr) == len(rg) == 8
        for j, (x, y) in enumerate(zip(rr, rg)):
            assert abs(float(x) - float(y)) <= 5e-6 * max(np.abs(cols[:, j]).max(), 1e-300), (rr, rg)


@pytest.mark.parametrize("mode", ["tasks", "fused"])
def test_contact_force_global_quantities_match_reference(mode):
    """"contactx/y" global quantities (GlobalQuantity.cpp:905-968): the force a plate of rigid contact particles exerts on a disk, by
    material and in total, averaged over the steps since the last global archive and cleared after reading.  The device sums the
    rigid field's force row over its active nodes (mpmgpu_contact_forces); the adapter keeps the reference's own step bookkeeping."""
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("oracle/_ref/NairnMPM or host/_build/NairnMPM_gpu not built")
    xml = (inputs.rigid_contact_plate(inputs.disks2d(analysis=10, vel=3000.0, vmax=11.0, gap=0.0, maxtime=0.6, archive_ms=0.3, extra_header=inputs.multimaterial(0, 0.3)))
           .replace("</MPMHeader>", "<GlobalArchiveTime units=\"ms\">0.02</GlobalArchiveTime><GlobalArchive type=\"contactx\" material=\"2\"/>"
                    "<GlobalArchive type=\"contacty\" material=\"2\"/><GlobalArchive type=\"contactx\"/><GlobalArchive type=\"Kinetic Energy\"/></MPMHeader>"))
    dref, out_ref = run(REF, xml, ("-np", "4"))
    dgpu, out_gpu = run(GPU, xml, ("-fused",) if mode == "fused" else ())
    assert "GPU TASKS" in out_gpu
    rows_r = [ln.split("\t") for ln in open(os.path.join(dref, "res/disks.global")).read().splitlines() if not ln.startswith("#")]
    rows_g = [ln.split("\t") for ln in open(os.path.join(dgpu, "res/disks.global")).read().splitlines() if not ln.startswith("#")]
    assert len(rows_r) == len(rows_g) and len(rows_r) >= 10, (len(rows_r), len(rows_g))
    cols = np.array([[float(x) for x in r] for r in rows_r])
    assert np.all(np.abs(cols[:, 1:4]).max(axis=0) > 0), "a contact-force column is zero throughout: the input does not exercise it"
    for rr, rg in zip(rows_r, rows_g):
        assert len(rr) == len(rg) == 5
        for j, (x, y) in enumerate(zip(rr, rg)):
            assert abs(float(x) - float(y)) <= 5e-6 * max(np.abs(cols[:, j]).max(), 1e-300), (rr, rg)
