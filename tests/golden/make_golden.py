"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref,
built by oracle/build_ref.sh from /root/reference) on the XML inputs of tests/inputs.py.

    python tests/golden/make_golden.py            # regenerates every *.npz here

Each fixture holds: the reference's state after set-up (particles p0, grid, BC list, materials),
node + particle dumps after EVERY task of step 1, and particle/node snapshots after N steps.
Only this script and the reference produce these files; the tests only read them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refharness import run_reference  # noqa: E402
from tests import inputs  # noqa: E402
from tests.parity import dedupe_per_task  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

DISK1 = '<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>'
DISK2 = '<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>'

CASES = {
    # name: (xml text, snapshots, per-task steps[, position jitter, velocity jitter])
    "block3d_jitter": (inputs.block3d(ncell=5, margin=3, E=100.0, vx=3.0e3, vy=-2.0e3, vz=-6.0e3), (1, 20, 60), 1, 0.35, 4000.0),
    "disks2d_ugimp_planestrain": (inputs.disks2d(analysis=10, gimp="uGIMP"), (1, 100), 1),
    # symmetry planes (config 1: the reference's TwoDisks input has them on all four grid edges): the left disk hits the plane
    # x = 0 (its mirror image), a second plane along y = -6 touches it from below; the BCs beside the planes reflect
    "disks2d_symmetry_planes": (inputs.disks2d(analysis=10, gimp="uGIMP", vel=3000.0, hmax=16.0, vmax=9.0)
                                .replace('<Horiz cellsize="1.0"/><Vert cellsize="1.0"/>', '<Horiz cellsize="1.0" symmax="0"/><Vert cellsize="1.0" symmin="-6"/>')
                                .replace('<Grid xmin="-16.0" xmax="16.0" ymin="-9.0"', '<Grid xmin="-16.0" xmax="2.0" ymin="-7.0"')
                                .split('<Body matname="Disk 2"')[0] + '</MaterialPoints>' +
                                inputs.disks2d(analysis=10, gimp="uGIMP", vel=3000.0).split('</MaterialPoints>')[1], (1, 2, 100), 2),
    "disks2d_linear_planestress": (inputs.disks2d(analysis=11, gimp=None, method=2), (1, 100), 1),
    "block3d_neohookean": (inputs.block3d(ncell=4, margin=3, material=inputs.neohookean_material(), vz=-8.0e3, vx=2.0e3), (1, 60), 1, 0.3, 2000.0),
    "block3d_neohookean_uj1": (inputs.block3d(ncell=3, margin=3, material=inputs.neohookean_material(ujoption=1), vz=-8.0e3), (1, 40), 1),
    "block3d_isoplastic": (inputs.block3d(ncell=4, margin=3, material=inputs.isoplastic_material(), vz=-4.0e4, vx=5.0e3), (1, 80), 1, 0.3, 3000.0),
    "disks2d_neohookean": (inputs.disks2d(analysis=10).replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>', '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha></Material>'), (1, 100), 1),
    "disks2d_isoplastic": (inputs.disks2d(analysis=10, vel=6000.0).replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>', '<Material Type="9" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep></Material>'), (1, 100), 1),
    # 2D: a disk hits a plate of rigid-BC particles that fixes x and moves in y
    "disks2d_rigid_plate": (inputs.disks2d(analysis=10, vel=4000.0)
                            .replace('<Body matname="Disk 2" angle="0" thick="1" vx="-4000.0" vy="0">\n      <Oval xmin="0.5" xmax="12.5" ymin="-6.0" ymax="6.0"/>',
                                     '<Body matname="Disk 2" angle="0" thick="1" vx="0" vy="300">\n      <Rect xmin="0" xmax="1" ymin="-7" ymax="7"/>')
                            .replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                     '<Material Type="11" Name="Disk 2"><SetDirection>3</SetDirection></Material>'), (1, 2, 100), 2),
    "disks2d_isoplastic_planestress": (inputs.disks2d(analysis=11, vel=6000.0).replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>', '<Material Type="9" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep></Material>'), (1, 100), 1),
    "block3d_xpic3": (inputs.block3d(ncell=4, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, custom_tasks=inputs.periodic_xpic(3, False, 2)), (1, 2, 3, 30), 2, 0.3, 3000.0),
    "block3d_fmpm2": (inputs.block3d(ncell=4, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, custom_tasks=inputs.periodic_xpic(2, True, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    "disks2d_fmpm3_neo": (inputs.disks2d(analysis=10, extra_header="").replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>', '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha></Material>').replace("</JANFEAInput>", inputs.periodic_xpic(3, True, 1) + "</JANFEAInput>"), (1, 2, 60), 2),
    "block3d_lcpdi_neo_xpic2": (inputs.block3d(ncell=3, margin=3, gimp="lCPDI", material=inputs.neohookean_material(), vz=-8.0e3, vx=2.0e3,
                                               custom_tasks=inputs.periodic_xpic(2, False, 1)), (1, 2, 40), 2, 0.3, 2000.0),
    "block3d_lcpdi_rcrit": (inputs.block3d(ncell=3, margin=3, gimp="lCPDI", E=50.0, vz=-2.0e4, vx=1.0e4, extra_header="<CPDIrcrit>0.6</CPDIrcrit>"), (1, 40), 1, 0.3, 5000.0),
    "disks2d_lcpdi": (inputs.disks2d(analysis=10, gimp="lCPDI"), (1, 100), 1),
    "disks2d_qcpdi": (inputs.disks2d(analysis=10, gimp="qCPDI"), (1, 100), 1),
    # small cases for the particle-update / post-extrapolation variants (PIC = XPIC(1), FMPM(1), USAVG- and USL- with and
    # without higher-order velocity extrapolation)
    "block3d_pic": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, custom_tasks=inputs.periodic_xpic(1, False, 1)), (1, 30), 1, 0.3, 3000.0),
    "block3d_fmpm1": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, custom_tasks=inputs.periodic_xpic(1, True, 1)), (1, 30), 1, 0.3, 3000.0),
    "block3d_usavg_minus": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, extra_header="<SkipPostExtrapolation/>"), (1, 30), 1, 0.3, 3000.0),
    "block3d_usavg_minus_xpic2": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, extra_header="<SkipPostExtrapolation/>",
                                                 custom_tasks=inputs.periodic_xpic(2, False, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    "block3d_usl_minus_fmpm2": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, method=3, extra_header="<SkipPostExtrapolation/>",
                                               custom_tasks=inputs.periodic_xpic(2, True, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    "block3d_usf_fmpm2": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, method=0, custom_tasks=inputs.periodic_xpic(2, True, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    "block3d_neohookean_av": (inputs.block3d(ncell=3, margin=3, material=inputs.neohookean_material(av=(0.3, 1.5)), vz=-3.0e4, vx=4.0e3), (1, 40), 1, 0.3, 3000.0),
    "block3d_isoplastic_av": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_material(av=(0.2, 2.0)), vz=-5.0e4), (1, 40), 1, 0.3, 3000.0),
    "block3d_rigid_wall": (inputs.block3d(ncell=4, margin=3, material=inputs.isoplastic_material(), vz=-4.0e4, bc=False,
                                          rigid=("wall", 4, (0.0, 0.0, 0.0))), (1, 2, 80), 2, 0.3, 2000.0),
    "block3d_rigid_wall_lattice": (inputs.block3d(ncell=2, margin=3, material=inputs.isoplastic_material(), vz=-4.0e4, bc=False,
                                                  rigid=("wall", 4, (0.0, 0.0, 0.0))), (1, 5), 0),
    "block3d_rigid_mirrored": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-8.0e3, vx=1.0e3, bc=False, rigid=("mirror_wall", 4, (0.0, 0.0, 0.0))), (1, 2, 40), 2, 0.3, 2000.0),
    "block3d_lcpdi_rigid_wall": (inputs.block3d(ncell=3, margin=3, gimp="lCPDI", E=100.0, vz=-8.0e3, vx=2.0e3, bc=False, rigid=("wall", 5, (0.0, 0.0, 0.0))), (1, 40), 1, 0.3, 2000.0),
    "block3d_rigid_piston": (inputs.block3d(ncell=4, margin=3, E=100.0, vz=0.0, rigid=("piston", 7, (1.5e3, -1.0e3, -8.0e3))), (1, 2, 60), 2, 0.3, 1500.0),
    "block3d_rigid_linear_xpic2": (inputs.block3d(ncell=3, margin=3, E=100.0, gimp=None, vz=-5.0e3, bc=False, rigid=("wall", 5, (0.0, 0.0, 0.0)),
                                                  custom_tasks=inputs.periodic_xpic(2, False, 1)), (1, 2, 40), 2, 0.3, 1500.0),
    # small-strain / large-rotation hypoelasticity (<largeRotation>1</largeRotation>): polar-decomposed strain increment,
    # stress and plastic strain rotated by dR (IsotropicMat::LRConstitutiveLaw, IsoPlasticity with useLargeRotation)
    "block3d_isotropic_lr": (inputs.block3d(ncell=4, margin=3, E=20.0, vz=-6.0e3, vx=5.0e3, vy=-2.0e3).replace("<alpha>0</alpha>", "<alpha>30</alpha><largeRotation>1</largeRotation>"),
                             (1, 80), 1, 0.3, 6000.0),
    "block3d_isoplastic_lr": (inputs.block3d(ncell=4, margin=3, material=inputs.isoplastic_material().replace("</Material>", "<largeRotation>1</largeRotation></Material>"),
                                             vz=-4.0e4, vx=1.5e4), (1, 80), 1, 0.3, 8000.0),
    "disks2d_lr_planestrain": (inputs.disks2d(analysis=10, vel=5000.0)
                               .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                        '<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha><largeRotation>1</largeRotation></Material>')
                               .replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                        '<Material Type="9" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep><largeRotation>1</largeRotation></Material>')
                               .replace('vx="-5000.0" vy="0"', 'vx="-5000.0" vy="1500"'), (1, 100), 1),
    "disks2d_lr_planestress": (inputs.disks2d(analysis=11, gimp=None, vel=5000.0)
                               .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                        '<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha><largeRotation>1</largeRotation></Material>')
                               .replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                        '<Material Type="9" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep><largeRotation>1</largeRotation></Material>')
                               .replace('vx="-5000.0" vy="0"', 'vx="-5000.0" vy="1500"'), (1, 100), 1),
    # branches of the laws no other case reaches: Neo-Hookean in plane stress (the three U(J) options solve for the zz stretch
    # differently, Neohookean.cpp:204-245), 2D artificial viscosity, softening with a minimum yield stress (LinearHardening
    # alphaMax), a material's own particle damping (<PDamping> inside <Material>, MaterialBase::GetMaterialDamping)
    "disks2d_neo_planestress": (inputs.disks2d(analysis=11, vel=4000.0)
                                .replace(DISK1, '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha><UJOption>1</UJOption></Material>')
                                .replace(DISK2, '<Material Type="28" Name="Disk 2"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha><UJOption>2</UJOption><PDamping>300</PDamping></Material>'),
                                (1, 100), 1),
    "disks2d_neo_planestress_av": (inputs.disks2d(analysis=11, vel=6000.0)
                                   .replace(DISK1, '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha><ArtificialVisc/><avA1>0.3</avA1><avA2>1.5</avA2></Material>'),
                                   (1, 100), 1),
    "block3d_isoplastic_softening": (inputs.block3d(ncell=4, margin=3, material=inputs.isoplastic_material(Ep=-1.0).replace("<Ep>-1.0</Ep>", "<Khard>-4.0</Khard><yieldMin>12.0</yieldMin>"),
                                                    vz=-6.0e4, vx=5.0e3), (1, 80), 1, 0.3, 3000.0),
    "block3d_material_pdamping": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3).replace("<alpha>0</alpha>", "<alpha>0</alpha><PDamping>2000</PDamping>"),
                                  (1, 30), 1, 0.3, 3000.0),
    # free-flying 3D blocks (no grid BCs): also run by the host-compiled device source, tests/test_device_step_cpu.py
    "block3d_free_ugimp": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-6.0e3, vx=3.0e3, bc=False), (1, 30), 1, 0.3, 4000.0),
    "block3d_free_lcpdi_xpic2": (inputs.block3d(ncell=3, margin=3, gimp="lCPDI", material=inputs.neohookean_material(), vz=-8.0e3, vx=2.0e3, bc=False,
                                                custom_tasks=inputs.periodic_xpic(2, False, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    # Mooney-Rivlin (MaterialID 8): 3D with artificial viscosity, plane strain / plane stress with the three U(J) options
    "block3d_mooney": (inputs.block3d(ncell=3, margin=3, material=inputs.mooney_material(av=(0.3, 1.5)), vz=-1.0e4, vx=3.0e3), (1, 40), 1, 0.3, 3000.0),
    "block3d_mooney_uj2": (inputs.block3d(ncell=3, margin=3, material=inputs.mooney_material(ujoption=2), vz=-1.0e4, bc=False), (1, 30), 1, 0.3, 3000.0),
    "disks2d_mooney_planestrain": (inputs.disks2d(analysis=10, vel=4000.0)
                                   .replace(DISK1, inputs.mooney_material(0.3, 0.1, 1.0, 1, name="Disk 1", rho=1.5))
                                   .replace(DISK2, inputs.mooney_material(0.25, 0.15, 1.0, None, name="Disk 2", rho=1.5)), (1, 100), 1),
    "disks2d_mooney_planestress": (inputs.disks2d(analysis=11, vel=4000.0)
                                   .replace(DISK1, inputs.mooney_material(0.3, 0.1, 1.0, 1, name="Disk 1", rho=1.5))
                                   .replace(DISK2, inputs.mooney_material(0.25, 0.15, 1.0, 2, name="Disk 2", rho=1.5)), (1, 100), 1),
    "disks2d_mooney_planestress_uj0": (inputs.disks2d(analysis=11, gimp=None, vel=5000.0)
                                       .replace(DISK1, inputs.mooney_material(0.3, 0.1, 1.0, 0, av=(0.2, 2.0), name="Disk 1", rho=1.5)), (1, 100), 1),
    # IsoPlasticity with the numerically returned hardening laws (bracketed Newton of HardeningLawBase): power laws, Johnson-Cook.
    # Short runs in 2D: the reference's solver stops at |d lambda/lambda| < 1e-4, so long runs of these soft, heavily yielding disks
    # amplify last-bit differences (tests/parity.py TOL_ITERATIVE); the law itself is pinned bit for bit by
    # tests/test_device_laws_vs_reference_cpu.py
    "block3d_isoplastic_nonlinear": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("Nonlinear"), vz=-4.0e4, vx=5.0e3), (1, 60), 1, 0.3, 3000.0),
    "block3d_isoplastic_nonlinear2_soft": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("Nonlinear2", Khard=-2.0, nhard=0.8, yieldMin=8.0),
                                                          vz=-5.0e4, vx=5.0e3), (1, 60), 1, 0.3, 3000.0),
    "block3d_johnsoncook": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("JohnsonCook", Djc=0.01), vz=-4.0e4, vx=5.0e3,
                                           extra_header="<StressFreeTemp>300</StressFreeTemp>"), (1, 60), 1, 0.3, 3000.0),
    "disks2d_johnsoncook_planestress": (inputs.disks2d(analysis=11, vel=3000.0, extra_header="<StressFreeTemp>300</StressFreeTemp>")
                                        .replace(DISK2, inputs.isoplastic_hardening_material("JohnsonCook", rho=1.5, E=1.0, yld=0.02, Bjc=0.03, name="Disk 2")), (1, 20), 1),
    "disks2d_nonlinear_planestrain_lr": (inputs.disks2d(analysis=10, vel=3000.0)
                                         .replace(DISK2, inputs.isoplastic_hardening_material("Nonlinear", rho=1.5, E=1.0, yld=0.02, name="Disk 2",
                                                                                              extra="<largeRotation>1</largeRotation>")), (1, 20), 1),
    "disks2d_nonlinear2_planestress": (inputs.disks2d(analysis=11, gimp=None, vel=3000.0)
                                       .replace(DISK1, inputs.isoplastic_hardening_material("Nonlinear2", rho=1.5, E=1.0, yld=0.02, name="Disk 1")), (1, 20), 1),
    # quadratic B-spline shape functions: B2SPLINE and its GIMP form B2GIMP (3D incl. rigid-BC particles, 2D)
    # SCGL hardening: shear modulus and yield stress follow the particle's pressure (and temperature: th3d_offset_scgl); yieldMax out of
    # reach (on the cap the reference's own return solver is decided by rounding noise: tests/test_device_laws_vs_reference_cpu.py)
    "block3d_scgl": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("SCGL", yieldMax=400.0, GPpG0=4.0e-4), vz=-4.0e4, vx=5.0e3,
                                    extra_header="<StressFreeTemp>300</StressFreeTemp>"), (1, 60), 1, 0.3, 3000.0),
    "disks2d_scgl_planestrain": (inputs.disks2d(analysis=10, vel=3000.0, extra_header="<StressFreeTemp>300</StressFreeTemp>")
                                 .replace(DISK2, inputs.isoplastic_hardening_material("SCGL", rho=1.5, E=1.0, yld=0.02, yieldMax=0.4, name="Disk 2")), (1, 100), 1),
    "block3d_b2spline": (inputs.block3d(ncell=3, margin=3, E=100.0, gimp="B2SPLINE", vz=-6.0e3, vx=3.0e3), (1, 30), 1, 0.3, 3000.0),
    "block3d_b2gimp": (inputs.block3d(ncell=3, margin=3, E=100.0, gimp="B2GIMP", vz=-6.0e3, vx=3.0e3, custom_tasks=inputs.periodic_xpic(2, True, 1)), (1, 2, 30), 2, 0.3, 3000.0),
    "block3d_b2gimp_rigid_wall": (inputs.block3d(ncell=3, margin=3, gimp="B2GIMP", material=inputs.isoplastic_material(), vz=-4.0e4, bc=False,
                                                 rigid=("wall", 4, (0.0, 0.0, 0.0))), (1, 40), 1, 0.3, 2000.0),
    "disks2d_b2spline": (inputs.disks2d(analysis=10, gimp="B2SPLINE"), (1, 60), 1),
    "disks2d_b2gimp_planestress": (inputs.disks2d(analysis=11, gimp="B2GIMP").replace(DISK1, '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha></Material>'),
                                   (1, 60), 1),
    "block3d_b2cpdi": (inputs.block3d(ncell=3, margin=3, gimp="B2CPDI", material=inputs.neohookean_material(), vz=-8.0e3, vx=2.0e3), (1, 30), 1, 0.3, 2000.0),
    "disks2d_b2cpdi": (inputs.disks2d(analysis=10, gimp="B2CPDI"), (1, 60), 1),
    # multimaterial mode: one velocity field per material + material contact.  The bodies touch
    # from the start (contact, and with it real strains, from the first step on) and the disks meet obliquely (inputs.oblique_disks: in the mirror-symmetric head-on case the friction law's tangent is rounding noise).
    "mm2d_friction_avgg": (inputs.oblique_disks(inputs.disks2d(analysis=10, vel=4000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(2, 0.3))), (1, 2, 60), 2),
    "mm2d_frictionless_maxg_position": (inputs.oblique_disks(inputs.disks2d(analysis=11, vel=4000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(0, None, 0.8))), (1, 2, 60), 2),
    "mm2d_stick_maxv_linear_usl": (inputs.oblique_disks(inputs.disks2d(analysis=10, gimp=None, method=3, vel=4000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(1, -1.0))), (1, 2, 60), 2),
    "mm2d_ignore_lcpdi_usf": (inputs.oblique_disks(inputs.disks2d(analysis=10, gimp="lCPDI", method=0, vel=4000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(2, -11.0))), (1, 2, 40), 2),
    "mm2d_friction_sn_powerlaw": (inputs.oblique_disks(inputs.disks2d(analysis=10, vel=4000.0, vmax=11.0, gap=0.0,
                                                       extra_header=inputs.multimaterial(4, 0.5, -0.7, ' Polar="0" Azimuth="8"'))), (1, 2, 60), 2),
    # (the reference itself crashes when a contact node is not handled as ONE PAIR that includes material field 0 -- three
    # materials on a node, or Normals="3" -- unless FMPM order > 1: MaterialContactNode::GetContactInfo indexes a NULL
    # FMPMContact array, MaterialContactNode.cpp:137-139; so two-material cases only)
    "mm3d_two_blocks_avgg_position": (inputs.blocks3d_contact(inputs.multimaterial(2, 0.4, 0.9), gimp=None, materials=2), (1, 2, 40), 2),
    "mm3d_two_blocks_maxg_stick_b2gimp": (inputs.blocks3d_contact(inputs.multimaterial(0, -1.0), gimp="B2GIMP", materials=2), (1, 2, 40), 2),
    "mm3d_two_blocks_maxv_friction_ugimp": (inputs.blocks3d_contact(inputs.multimaterial(1, 0.25), materials=2), (1, 2, 40), 2),
    # rigid contact materials (RigidMaterial with SetDirection 8): their particles keep their own field and velocity; every
    # other material of a node makes contact with it (RigidMaterialContactOnCVF)
    "mm2d_rigid_plate_maxg_friction": (inputs.rigid_contact_plate(inputs.disks2d(analysis=10, vel=3000.0, vmax=11.0, gap=0.0,
                                                                  extra_header=inputs.multimaterial(0, 0.3, None, ' RigidBias="10"'))), (1, 2, 60), 2),
    "mm3d_rigid_block_avgg_position_usl": (inputs.blocks3d_contact(inputs.multimaterial(2, None, 0.8), method=3, materials=2, rigid_b=True), (1, 2, 40), 2),
    "mm3d_rigid_block_maxv_stick_lcpdi": (inputs.blocks3d_contact(inputs.multimaterial(1, -1.0), gimp="lCPDI", materials=2, rigid_b=True), (1, 2, 40), 2),
    # heat conduction: bodies of different temperature, conductivity and heat capacity in contact;
    # one velocity field, then material velocity fields with frictional contact in 3D (transport stays on the node)
    "cond2d_disks_usavg": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=2000.0, vmax=11.0, gap=0.0, alpha=0.0)),
                                             (380.0, 290.0), (2000.0, 500.0), (800.0, 1500.0)), (1, 2, 60), 2),
    "cond2d_disks_lcpdi_usl_neo": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=11, gimp="lCPDI", method=3, vel=2000.0, vmax=11.0, gap=0.0, alpha=0.0))
                                                     .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>0.0</alpha>',
                                                              '<Material Type="28" Name="Disk 1"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>0</alpha>'),
                                                     (250.0, 420.0), (900.0, 3000.0), (1200.0, 700.0)), (1, 2, 60), 2),
    "cond3d_blocks_multimaterial": (inputs.conduction(inputs.blocks3d_contact(inputs.multimaterial(2, 0.3), materials=2).replace("<alpha>20</alpha>", "<alpha>0</alpha>"),
                                                      (400.0, 280.0), (5000.0, 1500.0), (600.0, 900.0)), (1, 2, 40), 2),
    # nodal temperature BCs: the bottom plane of the grid is held at 450 K and a strip on one side at 250 K (two BCs overlap on
    # the corner nodes); the block expands as it heats
    "cond3d_block_temperature_bcs": (inputs.conduction(inputs.block3d(ncell=4, margin=3, E=100.0, vz=-2.0e3, vx=1.0e3).replace("<alpha>0</alpha>", "<alpha>50</alpha>"),
                                                       (300.0,), (4000.0,), (700.0,))
                                     .replace("</GridBCs>", '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="4.01"><TempBC value="450"/></BCBox>'
                                                            '<BCBox xmin="-1" xmax="4.01" ymin="-1" ymax="20" zmin="-1" zmax="20"><TempBC value="250"/></BCBox></GridBCs>'),
                                     (1, 2, 40), 2, 0.3, 1500.0),
    # temperature BCs made by rigid particles (<SetTemperature/>): a hot piston pressing on a block whose bottom plane and one side carry
    # grid temperature BCs (the side's nodes under the piston stay with the grid BC), and a heater plate that controls nothing else
    "cond3d_rigid_hot_piston": (inputs.conduction(inputs.block3d(ncell=3, margin=3, E=100.0, vz=0.0, rigid=("piston", 4, (0.0, 0.0, -3.0e3)))
                                                  .replace("<SetDirection>4</SetDirection>", "<SetDirection>4</SetDirection><SetTemperature/>"),
                                                  (300.0, 600.0), (4000.0,), (700.0,))
                                .replace("</GridBCs>", '<BCBox xmin="-1" xmax="10" ymin="-1" ymax="10" zmin="-1" zmax="3.01"><TempBC value="350"/></BCBox>'
                                                       '<BCBox xmin="-1" xmax="3.01" ymin="-1" ymax="10" zmin="-1" zmax="10"><TempBC value="320"/></BCBox></GridBCs>'),
                                (1, 2, 40), 2, 0.3, 1000.0),
    "cond3d_rigid_heater_lcpdi_usl": (inputs.conduction(inputs.block3d(ncell=3, margin=3, E=100.0, gimp="lCPDI", method=3, vz=-2.0e3, vx=1.0e3, rigid=("wall", 0, (0.0, 0.0, 0.0)))
                                                        .replace("<SetDirection>0</SetDirection>", "<SetDirection>0</SetDirection><SetTemperature/>"),
                                                        (300.0, 500.0), (4000.0,), (700.0,)), (1, 2, 40), 2),
    # conduction under a mechanical FMPM(2) / XPIC(2) update (the transport update itself stays FLIP: no transport XPIC option)
    "cond3d_block_fmpm2_temperature_bcs": (inputs.conduction(inputs.block3d(ncell=3, margin=3, E=100.0, vz=-2.0e3, vx=1.0e3, custom_tasks=inputs.periodic_xpic(2, True, 1))
                                                             .replace("<alpha>0</alpha>", "<alpha>50</alpha>"), (300.0,), (4000.0,), (700.0,))
                                           .replace("</GridBCs>", '<BCBox xmin="-1" xmax="20" ymin="-1" ymax="20" zmin="-1" zmax="3.01"><TempBC value="450"/></BCBox></GridBCs>'),
                                           (1, 2, 40), 2, 0.3, 1500.0),
    "cond2d_disks_xpic2_usl": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=10, method=3, vel=2000.0, vmax=11.0, gap=0.0, alpha=60.0))
                                                 .replace("</JANFEAInput>", inputs.periodic_xpic(2, False, 1) + "</JANFEAInput>"),
                                                 (380.0, 290.0), (2000.0, 500.0), (800.0, 1500.0)), (1, 2, 40), 2),
    # particle heat-flux BCs (MatPtHeatFluxBC, external flux): heat fed into the top face of a moving block (uGIMP: undeformed corners)
    # and into one side of a disk (lCPDI, plane stress: deformed corners, thickness)
    "cond3d_heat_flux_ugimp": (inputs.particle_bcs(inputs.conduction(inputs.block3d(ncell=3, margin=3, E=100.0, vz=-2.0e3, vx=1.0e3), (300.0,), (4000.0,), (700.0,)), [
        ('<BCBox xmin="-1" xmax="10" ymin="-1" ymax="10" zmin="5.5" zmax="10">', '<HeatFluxBC dir="1" face="6" style="1" value="5e7"/>'),
        ('<BCBox xmin="5.5" xmax="10" ymin="-1" ymax="10" zmin="-1" zmax="10">', '<HeatFluxBC dir="1" face="2" style="1" value="-2e7"/>')]), (1, 2, 40), 2, 0.3, 1000.0),
    "cond2d_heat_flux_lcpdi_planestress": (inputs.particle_bcs(inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=11, gimp="lCPDI", vel=2000.0, vmax=11.0, gap=0.0, alpha=0.0)),
                                                                                 (380.0, 290.0), (2000.0, 500.0), (800.0, 1500.0)), [
        ('<BCLine x1="-12" y1="-11" x2="-12" y2="11" tolerance="3">', '<HeatFluxBC dir="1" face="4" style="1" value="3e5"/>')]), (1, 2, 40), 2),
    # thermal strains in the laws: conduction with expanding materials, and bodies that start off the stress-free temperature
    # (one temperature jump handed to the laws by the first particle update) -- every law and analysis type that carries the terms
    "th2d_cond_iso_planestrain": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=2000.0, vmax=11.0, gap=0.0, alpha=60.0)),
                                                    (380.0, 290.0), (2000.0, 500.0), (800.0, 1500.0)), (1, 2, 40), 2),
    "th2d_cond_isoplastic_neo_planestress": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=11, vel=3000.0, vmax=11.0, gap=0.0, alpha=60.0))
                                                               .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha>',
                                                                        '<Material Type="9" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep>')
                                                               .replace('<Material Type="1" Name="Disk 2"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha>',
                                                                        '<Material Type="28" Name="Disk 2"><rho>1.5</rho><G>0.4</G><K>1.0</K><alpha>60</alpha>'),
                                                               (390.0, 280.0), (2000.0, 500.0), (800.0, 1500.0)), (1, 2, 40), 2),
    "th3d_offset_iso": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-3.0e3, vx=1.0e3, extra_header="<StressFreeTemp>300</StressFreeTemp>")
                        .replace("<alpha>0</alpha>", "<alpha>80</alpha>").replace('<Body ', '<Body temp="340" ', 1), (1, 2, 40), 2, 0.3, 2000.0),
    "th3d_offset_isoplastic_usl": (inputs.block3d(ncell=3, margin=3, method=3, material=inputs.isoplastic_material(yld=5.0), vz=-2.0e4,
                                                  extra_header="<StressFreeTemp>300</StressFreeTemp>").replace('<Body ', '<Body temp="420" ', 1), (1, 2, 40), 2, 0.3, 2000.0),
    "th3d_offset_scgl": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("SCGL", yieldMax=400.0, GPpG0=4.0e-4, GTpG0=-1.0e-3), vz=-4.0e4,
                                        vx=5.0e3, extra_header="<StressFreeTemp>300</StressFreeTemp>").replace('<Body ', '<Body temp="380" ', 1), (1, 2, 40), 2, 0.3, 3000.0),
    "th3d_offset_neohookean": (inputs.block3d(ncell=3, margin=3, material=inputs.neohookean_material(), vz=-6.0e3, extra_header="<StressFreeTemp>300</StressFreeTemp>")
                               .replace('<Body ', '<Body temp="260" ', 1), (1, 2, 40), 2, 0.3, 2000.0),
    "th2d_offset_mooney_iso_planestress": (inputs.oblique_disks(inputs.disks2d(analysis=11, vel=3000.0, vmax=11.0, gap=0.0, alpha=60.0, extra_header="<StressFreeTemp>300</StressFreeTemp>"))
                                           .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha></Material>',
                                                    inputs.mooney_material(0.3, 0.1, 1.0, 1, name="Disk 1", rho=1.5))
                                           .replace('<Body matname="Disk 1"', '<Body temp="350" matname="Disk 1"').replace('<Body matname="Disk 2"', '<Body temp="270" matname="Disk 2"'),
                                           (1, 2, 40), 2),
    # <EnergyCoupling/>: adiabatic mode -- plastic work and the thermoelastic effect heat the particles (Johnson-Cook softening
    # feels it), without a transport task and together with conduction
    "th3d_adiabatic_johnsoncook": (inputs.block3d(ncell=3, margin=3, material=inputs.isoplastic_hardening_material("JohnsonCook", Djc=0.01), vz=-6.0e4, vx=5.0e3,
                                                  extra_header="<StressFreeTemp>300</StressFreeTemp>").replace("</JANFEAInput>", "<Thermal><EnergyCoupling/></Thermal></JANFEAInput>"),
                                   (1, 2, 60), 2, 0.3, 3000.0),
    "th2d_adiabatic_conduction_isoplastic": (inputs.conduction(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=4000.0, vmax=11.0, gap=0.0, alpha=60.0))
                                                               .replace('<Material Type="1" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60.0</alpha>',
                                                                        '<Material Type="9" Name="Disk 1"><rho>1.5</rho><E>1.0</E><nu>0.33</nu><alpha>60</alpha><Hardening>Linear</Hardening><yield>0.02</yield><Ep>0.1</Ep>'),
                                                               (320.0, 290.0), (2000.0, 500.0), (800.0, 1500.0)).replace("<Conduction/>", "<Conduction/><EnergyCoupling/>"),
                                             (1, 2, 40), 2),
    # particle traction BCs (MatPtTractionBC): pressure on the top face (normal, follows the deformed face), a shear stress on the +x
    # face (fixed direction) and a second BC on the corner particles; uGIMP (undeformed corners, deformed weight) and lCPDI (deformed)
    "trac3d_pressure_shear_ugimp": (inputs.particle_bcs(inputs.block3d(ncell=3, margin=3, E=100.0, vz=-2.0e3, vx=1.0e3), [
        ('<BCBox xmin="-1" xmax="10" ymin="-1" ymax="10" zmin="5.5" zmax="10">', '<TractionBC dir="11" face="6" style="1" stress="-8"/>'),
        ('<BCBox xmin="5.5" xmax="10" ymin="-1" ymax="10" zmin="-1" zmax="10">', '<TractionBC dir="3" face="2" style="1" stress="3"/>'),
        ('<BCBox xmin="-1" xmax="3.5" ymin="-1" ymax="3.5" zmin="-1" zmax="10">', '<TractionBC dir="1" face="4" style="1" stress="-2"/>')]), (1, 2, 40), 2, 0.3, 1000.0),
    "trac3d_pressure_lcpdi_usl": (inputs.particle_bcs(inputs.block3d(ncell=3, margin=3, E=100.0, gimp="lCPDI", method=3, vz=-2.0e3, vx=1.0e3), [
        ('<BCBox xmin="-1" xmax="10" ymin="-1" ymax="10" zmin="5.5" zmax="10">', '<TractionBC dir="11" face="6" style="1" stress="-8"/>'),
        ('<BCBox xmin="-1" xmax="10" ymin="5.5" ymax="10" zmin="-1" zmax="10">', '<TractionBC dir="2" face="3" style="1" stress="-4"/>')]), (1, 2, 40), 2),
    "trac2d_disks_normal_tangent": (inputs.particle_bcs(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=1000.0, vmax=11.0, gap=0.0)), [
        ('<BCLine x1="-12" y1="-11" x2="-12" y2="11" tolerance="3">', '<TractionBC dir="11" face="4" style="1" stress="-0.005"/>'),
        ('<BCLine x1="-20" y1="6" x2="20" y2="6" tolerance="2">', '<TractionBC dir="12" face="3" style="1" stress="0.003"/>'),
        ('<BCLine x1="12" y1="-11" x2="12" y2="11" tolerance="3">', '<TractionBC dir="1" face="2" style="1" stress="-0.004"/>')]), (1, 2, 40), 2),
    "trac2d_multimaterial_lcpdi_planestress": (inputs.particle_bcs(inputs.oblique_disks(inputs.disks2d(analysis=11, gimp="lCPDI", vel=2000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(2, 0.3))), [
        ('<BCLine x1="-12" y1="-11" x2="-12" y2="11" tolerance="3">', '<TractionBC dir="11" face="4" style="1" stress="-0.05"/>'),
        ('<BCLine x1="0" y1="-11" x2="0" y2="11" tolerance="3">', '<TractionBC dir="2" face="1" style="1" stress="0.03"/>')]), (1, 2, 40), 2),
    "trac3d_pressure_b2gimp": (inputs.particle_bcs(inputs.block3d(ncell=3, margin=3, E=100.0, gimp="B2GIMP", vz=-2.0e3, vx=1.0e3), [
        ('<BCBox xmin="-1" xmax="10" ymin="-1" ymax="10" zmin="5.5" zmax="10">', '<TractionBC dir="11" face="6" style="1" stress="-8"/>'),
        ('<BCBox xmin="5.5" xmax="10" ymin="-1" ymax="10" zmin="-1" zmax="10">', '<TractionBC dir="3" face="2" style="1" stress="3"/>')]), (1, 2, 30), 2, 0.3, 1000.0),
    "trac2d_disks_b2cpdi": (inputs.particle_bcs(inputs.oblique_disks(inputs.disks2d(analysis=10, gimp="B2CPDI", vel=1000.0, vmax=11.0, gap=0.0)), [
        ('<BCLine x1="-12" y1="-11" x2="-12" y2="11" tolerance="3">', '<TractionBC dir="11" face="4" style="1" stress="-0.005"/>'),
        ('<BCLine x1="-20" y1="6" x2="20" y2="6" tolerance="2">', '<TractionBC dir="12" face="3" style="1" stress="0.003"/>')]), (1, 2, 30), 2),
    # reaction forces of the velocity BCs (NodalVelBC::freaction summed by bcID: "reaction<step>" beside "reaction_ids")
    "react3d_walls_ugimp": (inputs.reaction_walls3d(E=100.0, vz=-3.0e3, vx=2.0e3, vy=-1.0e3, gravity=(0.0, 0.0, -5.0e5)), (1, 2, 40), 2, 0.3, 1000.0),
    "react3d_walls_lcpdi_usl": (inputs.reaction_walls3d(ncell=3, E=100.0, gimp="lCPDI", method=3, vz=-3.0e3, vx=2.0e3, vy=-1.0e3), (1, 2, 30), 2),
    "react3d_rigid_piston_fmpm2": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=0.0, rigid=("piston", 7, (1.0e3, -5.0e2, -6.0e3)),
                                                  custom_tasks=inputs.periodic_xpic(2, True, 1)), (1, 2, 30), 2, 0.3, 1000.0),
    "react3d_rigid_wall_xpic2": (inputs.block3d(ncell=3, margin=3, E=100.0, vz=-5.0e3, vx=1.0e3, bc=False, rigid=("wall", 4, (0.0, 0.0, 0.0)),
                                                custom_tasks=inputs.periodic_xpic(2, False, 1)), (1, 2, 30), 2, 0.3, 1000.0),
    "react2d_multimaterial_wall": (inputs.grid_bcs(inputs.oblique_disks(inputs.disks2d(analysis=10, vel=4000.0, vmax=11.0, gap=0.0, extra_header=inputs.multimaterial(2, 0.3))),
                                                   [('<BCLine x1="0" y1="-11" x2="0" y2="11" tolerance="1.1">', 'dir="1" vel="0" id="-4"')]), (1, 2, 40), 2),
    "block3d_ugimp_usavg": (inputs.block3d(ncell=4, margin=2), (1, 10, 100), 1),
    "block3d_fast_crossings": (inputs.block3d(ncell=4, margin=3, E=10.0, vx=2.0e4, vy=1.0e4, vz=-1.5e4), (1, 40), 1),
    "block3d_gravity_damping": (inputs.block3d(ncell=3, margin=2, vz=-500.0, gravity=(0.0, 0.0, -9.8e6),
                                               damping=50.0, pdamping=20.0), (1, 50), 1),
    "block3d_linear_usl": (inputs.block3d(ncell=3, margin=2, gimp=None, method=3), (1, 50), 1),
    "block3d_ugimp_usf": (inputs.block3d(ncell=3, margin=2, method=0), (1, 50), 1),
}

KEEP_P = ("pos", "vel", "sp", "pressure", "ep", "wrot", "eplast", "energies", "hist", "inElem", "crossings", "ncpos", "acc", "temperature")


def slim(z):
    out = {}
    for k, v in z.items():
        parts = k.split("/")
        if parts[0].startswith("s") and parts[1].startswith("t") and parts[2] == "p":
            if parts[3] not in KEEP_P:
                continue
        out[k] = v
    return out


def main(names=None):
    for name, spec in CASES.items():
        if names and name not in names:
            continue
        xml, snaps, pts = spec[:3]
        ja, va = (spec[3], spec[4]) if len(spec) > 3 else (0.0, 0.0)
        z = run_reference(xml, snaps=snaps, per_task_steps=pts, nprocs=1, jitter_amp=ja, vel_amp=va)
        z = slim(z)
        z = dedupe_per_task(z)          # per-task entries identical to the previous task's are left out (tests/parity.py puts them back)
        z["xml"] = np.array(xml)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **z)
        print(name, "%.1f KB" % (os.path.getsize(path) / 1024.0), "tasks:", list(z["task_names"]))


if __name__ == "__main__":
    main(sys.argv[1:])
