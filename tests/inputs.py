"""XML inputs (NairnMPM.dtd format) for the parity tests and the bench.

These are the synthetic cases SURVEY.md section 8(d) names, written in the reference's own input
format so the same text drives the reference (oracle/_ref: goldens, CPU baseline) and the drop-in driver
(host/_build/NairnMPM_gpu); the Python host gets the same problems from the reference's own set-up dump
(problem.from_reference_dump) or from its generators (problem.block3d).
"""


def block3d(ncell=4, margin=2, E=1000.0, nu=0.3, rho=1.0, vz=-1000.0, vx=0.0, vy=0.0, cfl=0.4,
            method=2, gimp="uGIMP", maxtime=1.0, material=None, extra_header="", bc=True,
            gravity=None, damping=None, pdamping=None, ppc=None, custom_tasks="", rigid=None):
    """3D block of ncell^3 cells (8 particles per cell) inside a (ncell+2*margin)^3 grid, 1 mm cells.

    Config 2 of BASELINE.json is block3d(ncell=50, margin=7).  The bottom plane z<=margin is held
    with a zero z-velocity grid BC.

    rigid = (where, set_direction, (vx, vy, vz)) adds a one-cell-thick plate of rigid-BC particles
    (RigidMaterial, Type 11) under ("wall") or on top of ("piston") the block, one cell wider than it
    (BASELINE config 4 family: Taylor bar against a rigid wall).  A fourth entry lists setting functions of time (ms) and
    position for the controlled directions (RigidMaterial SettingFunction, SettingFunction2, ...).
    """
    n = ncell + 2 * margin
    lo, hi = margin, margin + ncell
    mat = material or ('<Material Type="1" Name="Blk"><rho>%r</rho><E>%r</E><nu>%r</nu><alpha>0</alpha></Material>'
                       % (rho, E, nu))
    gimp_tag = '<GIMP type="%s"/>' % gimp if gimp else ""
    ppc_tag = "<MatlPtsPerElement>%d</MatlPtsPerElement>" % ppc if ppc else ""
    bcs = ""
    if bc:
        bcs = ('<GridBCs><BCBox xmin="-1" xmax="%d" ymin="-1" ymax="%d" zmin="-1" zmax="%g">'
               '<DisBC dir="3" vel="0"/></BCBox></GridBCs>' % (n + 1, n + 1, lo + 0.01))
    grav = ""
    if gravity:
        grav = '<Gravity><BodyXForce>%r</BodyXForce><BodyYForce>%r</BodyYForce><BodyZForce>%r</BodyZForce></Gravity>' % tuple(gravity)
    damp = ""
    if damping is not None:
        damp += "<Damping>%r</Damping>" % damping
    if pdamping is not None:
        damp += "<PDamping>%r</PDamping>" % pdamping
    rigid_body = ""
    if rigid:
        where, setdir, rv = rigid[:3]
        mirrored = 0
        if where == "mirror_wall":      # two cells thick, overlapping the first cell layer of the block, reflecting at its lower edge
            z0, z1, mirrored = lo - 1, lo + 1, -1
        else:
            z0, z1 = (lo - 1, lo) if where == "wall" else (hi, hi + 1)
        rigid_body = ('<Body matname="Plate" vx="%r" vy="%r" vz="%r"><Box xmin="%d" xmax="%d" ymin="%d" ymax="%d" zmin="%d" zmax="%d"/></Body>'
                      % (rv[0], rv[1], rv[2], lo - 1, hi + 1, lo - 1, hi + 1, z0, z1))
        fxn = "".join("<SettingFunction%s>%s</SettingFunction%s>" % ("" if i == 0 else str(i + 1), f, "" if i == 0 else str(i + 1))
                      for i, f in enumerate(rigid[3])) if len(rigid) > 3 else ""
        if mirrored:
            fxn += "<mirrored>%d</mirrored>" % mirrored
        mat += '<Material Type="11" Name="Plate"><SetDirection>%d</SetDirection>%s</Material>' % (setdir, fxn)
    return """<?xml version='1.0'?>
<!DOCTYPE JANFEAInput SYSTEM "NairnMPM.dtd">
<JANFEAInput version='3'>
  <Header><Description>3D block</Description><Analysis>12</Analysis></Header>
  <MPMHeader>
    <MPMMethod>%d</MPMMethod>
    <Timing step="1e-3" max="%r" CFL="%r" units="ms"/>
    <ArchiveTime units="ms">1000</ArchiveTime>
    <ArchiveRoot>res/blk.</ArchiveRoot>
    <MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>
    %s %s %s %s
  </MPMHeader>
  <Mesh output="file">
    <Grid xmin="0" xmax="%d" ymin="0" ymax="%d" zmin="0" zmax="%d">
      <Horiz cellsize="1"/><Vert cellsize="1"/><Depth cellsize="1"/>
    </Grid>
  </Mesh>
  <MaterialPoints>
    <Body matname="Blk" vx="%r" vy="%r" vz="%r">
      <Box xmin="%d" xmax="%d" ymin="%d" ymax="%d" zmin="%d" zmax="%d"/>
    </Body>
    %s
  </MaterialPoints>
  %s
  %s
  %s
  %s
</JANFEAInput>
""" % (method, maxtime, cfl, gimp_tag, ppc_tag, damp, extra_header, n, n, n, vx, vy, vz,
       lo, hi, lo, hi, lo, hi, rigid_body, mat, bcs, grav, custom_tasks)


def disks2d(analysis=10, gimp="uGIMP", method=2, cell=1.0, radius=6.0, gap=1.0, hmax=16.0, vmax=9.0, E=1.0, nu=0.33,
            rho=1.5, vel=2500.0, alpha=60.0, maxtime=1.0, archive_ms=1000.0, root="res/disks.", extra_header=""):
    """2D two-disk head-on impact (BASELINE config 1 family; modelled on the reference's TwoDisks example but
    without symmetry planes): analysis 10 = plane strain, 11 = plane stress; gimp 'uGIMP' or None (Classic)."""
    gimp_tag = '<GIMP type="%s"/>' % gimp if gimp else ""
    x1 = -(gap / 2.0 + 2.0 * radius)
    return """<?xml version='1.0'?>
<!DOCTYPE JANFEAInput SYSTEM "NairnMPM.dtd">
<JANFEAInput version='3'>
  <Header><Description>2D disks</Description><Analysis>%d</Analysis></Header>
  <MPMHeader>
    <MPMMethod>%d</MPMMethod>
    <MaxTime units="ms">%r</MaxTime>
    <ArchiveTime units="ms">%r</ArchiveTime>
    <ArchiveRoot>%s</ArchiveRoot>
    <MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>
    %s %s
  </MPMHeader>
  <Mesh output="file">
    <Grid xmin="-%r" xmax="%r" ymin="-%r" ymax="%r">
      <Horiz cellsize="%r"/><Vert cellsize="%r"/>
    </Grid>
  </Mesh>
  <MaterialPoints>
    <Body matname="Disk 1" angle="0" thick="1" vx="%r" vy="0">
      <Oval xmin="%r" xmax="%r" ymin="-%r" ymax="%r"/>
    </Body>
    <Body matname="Disk 2" angle="0" thick="1" vx="-%r" vy="0">
      <Oval xmin="%r" xmax="%r" ymin="-%r" ymax="%r"/>
    </Body>
  </MaterialPoints>
  <Material Type="1" Name="Disk 1"><rho>%r</rho><E>%r</E><nu>%r</nu><alpha>%r</alpha></Material>
  <Material Type="1" Name="Disk 2"><rho>%r</rho><E>%r</E><nu>%r</nu><alpha>%r</alpha></Material>
</JANFEAInput>
""" % (analysis, method, maxtime, archive_ms, root, gimp_tag, extra_header, hmax, hmax, vmax, vmax, cell, cell,
       vel, x1, x1 + 2 * radius, radius, radius, vel, -x1 - 2 * radius, -x1, radius, radius,
       rho, E, nu, alpha, rho, E, nu, alpha)


NEOHOOKEAN_MAT = ('<Material Type="28" Name="Blk"><rho>1.0</rho><G>%r</G><K>%r</K><alpha>40</alpha>%s</Material>')
ISOPLASTIC_MAT = ('<Material Type="9" Name="Blk"><rho>%r</rho><E>%r</E><nu>%r</nu><alpha>20</alpha>'
                  '<Hardening>Linear</Hardening><yield>%r</yield><Ep>%r</Ep></Material>')


AV_TAGS = "<ArtificialVisc/><avA1>%r</avA1><avA2>%r</avA2>"


def neohookean_material(G=40.0, K=200.0, ujoption=None, av=None):
    extra = "" if ujoption is None else "<UJOption>%d</UJOption>" % ujoption
    if av is not None:
        extra += AV_TAGS % av
    return NEOHOOKEAN_MAT % (G, K, extra)


def mooney_material(G1=30.0, G2=10.0, K=200.0, ujoption=None, av=None, name="Blk", rho=1.0):
    extra = "" if ujoption is None else "<UJOption>%d</UJOption>" % ujoption
    if av is not None:
        extra += AV_TAGS % av
    return '<Material Type="8" Name="%s"><rho>%r</rho><G1>%r</G1><G2>%r</G2><K>%r</K><alpha>40</alpha>%s</Material>' % (name, rho, G1, G2, K, extra)


def isoplastic_material(rho=2.0, E=2000.0, nu=0.33, yld=20.0, Ep=100.0, av=None):
    m = ISOPLASTIC_MAT % (rho, E, nu, yld, Ep)
    if av is not None:
        m = m.replace("</Material>", AV_TAGS % av + "</Material>")
    return m


def isoplastic_hardening_material(law, rho=2.0, E=2000.0, nu=0.33, yld=20.0, name="Blk", **k):
    """IsoPlasticity with <Hardening>Nonlinear|Nonlinear2|JohnsonCook|SCGL</Hardening> and that law's own properties."""
    if law == "SCGL":       # SCGLHardening.cpp:38-69: pressure- and temperature-dependent shear modulus, capped power-law yield
        props = ("<yield>%r</yield><betahard>%r</betahard><nhard>%r</nhard><yieldMax>%r</yieldMax><GPpG0>%r</GPpG0><GTpG0>%r</GTpG0>"
                 % (yld, k.get("betahard", 8.0), k.get("nhard", 0.4), k.get("yieldMax", 1.6 * yld), k.get("GPpG0", 0.5 / E), k.get("GTpG0", -2.2e-4)))
    elif law in ("Nonlinear", "Nonlinear2"):
        props = "<yield>%r</yield><Khard>%r</Khard><nhard>%r</nhard>" % (yld, k.get("Khard", 8.0), k.get("nhard", 0.4))
        if k.get("yieldMin") is not None:
            props += "<yieldMin>%r</yieldMin>" % k["yieldMin"]
    else:
        props = ("<Ajc>%r</Ajc><Bjc>%r</Bjc><njc>%r</njc><Cjc>%r</Cjc><ep0jc>%r</ep0jc><Tmjc>%r</Tmjc><mjc>%r</mjc>"
                 % (yld, k.get("Bjc", 30.0), k.get("njc", 0.5), k.get("Cjc", 0.02), k.get("ep0jc", 1.0), k.get("Tmjc", 1600.0), k.get("mjc", 1.1)))
        if k.get("Djc"):
            props += "<Djc>%r</Djc><n2jc>%r</n2jc>" % (k["Djc"], k.get("n2jc", 2.0))
    extra = k.get("extra", "")
    return ('<Material Type="9" Name="%s"><rho>%r</rho><E>%r</E><nu>%r</nu><alpha>20</alpha><Hardening>%s</Hardening>%s%s</Material>'
            % (name, rho, E, nu, law, props, extra))


def periodic_xpic(order, fmpm=False, periodic_steps=1):
    """<CustomTasks> block scheduling the reference's PeriodicXPIC task (Custom_Tasks/PeriodicXPIC.cpp:62-157)."""
    return ('<CustomTasks><Schedule name="PeriodicXPIC"><Parameter name="%s">%d</Parameter>'
            '<Parameter name="periodicSteps">%d</Parameter></Schedule></CustomTasks>'
            % ("FMPMOrder" if fmpm else "XPICOrder", order, periodic_steps))


def oblique_disks(xml, vy=1300.0, shift=2.5):
    """Break the mirror symmetry of disks2d: the second disk is shifted in y and moves in y too, so the contact between
    the disks is oblique.  (In the head-on case the tangential stick momentum is rounding noise, which the reference's
    friction law normalises into a direction: CoulombFriction.cpp:217-229.)"""
    a, b = xml.split('<Body matname="Disk 2"')
    b = b.replace('vy="0"', 'vy="%r"' % vy, 1).replace('ymin="-6.0" ymax="6.0"', 'ymin="%r" ymax="%r"' % (-6.0 + shift, 6.0 + shift), 1)
    return a + '<Body matname="Disk 2"' + b


def multimaterial(normals=2, friction=None, position=None, extra=""):
    """<MultiMaterialMode> header element: Normals 0 MAXG, 1 MAXV, 2 AVGG, 3 OWNG, 4 SN (MPMReadHandler.cpp:607-652);
    friction: None (default law: frictionless), < -10 ignore, < 0 stick, else the coefficient; position: <ContactPosition>."""
    inner = ""
    if friction is not None:
        inner += "<Friction>%r</Friction>" % friction
    if position is not None:
        inner += "<ContactPosition>%r</ContactPosition>" % position
    return '<MultiMaterialMode Normals="%d"%s>%s</MultiMaterialMode>' % (normals, extra, inner)


def conduction(xml, temps, kcond, cp, stress_free=300.0):
    """Switch heat conduction on in an input: <Thermal><Conduction/></Thermal>, a start temperature per <Body> (temps), kCond and
    Cp per <Material> (in input order; materials must have zero thermal expansion on the device path), <StressFreeTemp>."""
    parts = xml.split("<Body ")
    assert len(parts) - 1 == len(temps), (len(parts) - 1, temps)
    xml = parts[0] + "".join('<Body temp="%r" %s' % (t, rest) for t, rest in zip(temps, parts[1:]))
    parts = xml.split("</Material>")
    assert len(parts) - 1 >= len(kcond)
    xml = "".join(seg + ("<kCond>%r</kCond><Cp>%r</Cp></Material>" % (kcond[i], cp[i]) if i < len(kcond) else ("</Material>" if i < len(parts) - 1 else ""))
                  for i, seg in enumerate(parts))
    xml = xml.replace("</MPMHeader>", "<StressFreeTemp>%r</StressFreeTemp></MPMHeader>" % stress_free)
    return xml.replace("</JANFEAInput>", "<Thermal><Conduction/></Thermal></JANFEAInput>")


def rigid_contact_plate(xml):
    """disks2d with the second disk turned into a plate of RIGID CONTACT particles (RigidMaterial, SetDirection 8) that moves
    against the first disk at an angle."""
    a, b = xml.split('<Body matname="Disk 2"')
    vx = b.split('vx="')[1].split('"')[0]
    b = b.replace('vx="%s" vy="0"' % vx, 'vx="-2000.0" vy="700"', 1).replace('<Oval xmin="0.0" xmax="12.0" ymin="-6.0" ymax="6.0"/>', '<Rect xmin="0.0" xmax="2.0" ymin="-8.0" ymax="8.0"/>', 1)
    xml = a + '<Body matname="Disk 2"' + b
    head, tail = xml.split('<Material Type="1" Name="Disk 2">')
    return head + '<Material Type="11" Name="Disk 2"><SetDirection>8</SetDirection></Material>' + tail.split("</Material>", 1)[1]


def blocks3d_contact(header, gimp="uGIMP", method=2, materials=3, rigid_b=False):
    """Three (or two) 3D blocks of different materials flying into each other inside a 12 x 10 x 10 grid (A and B touch from the start, so the first steps already carry contact): nodes seen by two
    and by three materials (the lumped branch of MaterialContactOnCVFLumped)."""
    gimp_tag = '<GIMP type="%s"/>' % gimp if gimp else ""
    third = ('<Body matname="C" vx="-500" vy="-4000" vz="-1500"><Box xmin="4" xmax="7" ymin="6.5" ymax="8.5" zmin="3.5" zmax="6.5"/></Body>'
             if materials > 2 else "")
    third_mat = ('<Material Type="28" Name="C"><rho>1.2</rho><G>30</G><K>90</K><alpha>40</alpha></Material>' if materials > 2 else "")
    return """<?xml version='1.0'?>
<!DOCTYPE JANFEAInput SYSTEM "NairnMPM.dtd">
<JANFEAInput version='3'>
  <Header><Description>3D blocks in contact</Description><Analysis>12</Analysis></Header>
  <MPMHeader>
    <MPMMethod>%d</MPMMethod>
    <Timing step="1e-3" max="1.0" CFL="0.4" units="ms"/>
    <ArchiveTime units="ms">1000</ArchiveTime>
    <ArchiveRoot>res/blk.</ArchiveRoot>
    <MPMArchiveOrder>iYYYYNNNNNNNYNNNNY</MPMArchiveOrder>
    %s %s
  </MPMHeader>
  <Mesh output="file">
    <Grid xmin="0" xmax="12" ymin="0" ymax="10" zmin="0" zmax="10">
      <Horiz cellsize="1"/><Vert cellsize="1"/><Depth cellsize="1"/>
    </Grid>
  </Mesh>
  <MaterialPoints>
    <Body matname="A" vx="5000" vy="600" vz="-300"><Box xmin="2" xmax="5" ymin="3" ymax="6" zmin="3" zmax="6"/></Body>
    <Body matname="B" vx="-4000" vy="-200" vz="900"><Box xmin="5" xmax="8" ymin="3.5" ymax="6.5" zmin="4" zmax="7"/></Body>
    %s
  </MaterialPoints>
  <Material Type="1" Name="A"><rho>1.0</rho><E>100</E><nu>0.3</nu><alpha>0</alpha></Material>
  <Material Type="9" Name="B"><rho>2.0</rho><E>400</E><nu>0.33</nu><alpha>20</alpha><Hardening>Linear</Hardening><yield>8</yield><Ep>40</Ep></Material>
  %s
</JANFEAInput>
""".replace('<Material Type="9" Name="B"><rho>2.0</rho><E>400</E><nu>0.33</nu><alpha>20</alpha><Hardening>Linear</Hardening><yield>8</yield><Ep>40</Ep></Material>',
             '<Material Type="11" Name="B"><SetDirection>8</SetDirection></Material>' if rigid_b else
             '<Material Type="9" Name="B"><rho>2.0</rho><E>400</E><nu>0.33</nu><alpha>20</alpha><Hardening>Linear</Hardening><yield>8</yield><Ep>40</Ep></Material>') % (method, gimp_tag, header, third, third_mat)


def grid_bcs(xml, blocks):
    """Append a <GridBCs> element made of (shape element, DisBC attributes) pairs, e.g.
    ('<BCBox xmin="0" .../>'-style opening tag without the slash, 'dir="3" vel="0" id="-1"')."""
    body = "".join("%s<DisBC %s/></%s>" % (shape, bc, shape[1:].split()[0].rstrip(">")) for shape, bc in blocks)
    return xml.replace("</MaterialPoints>", "</MaterialPoints><GridBCs>%s</GridBCs>" % body, 1)


def reaction_walls3d(ncell=4, margin=3, **kw):
    """block3d held by four sets of velocity BCs with their own ids (the "reactionx/y/z" global quantities sum NodalVelBC::freaction
    by id): the bottom plane (z, id -1), the +x face the block moves into (x, id -2), the top plane pushed down at constant
    velocity (z, id -3) and the -y face with a skewed xy condition (id -4)."""
    n, lo, hi = ncell + 2 * margin, margin, margin + ncell
    box = '<BCBox xmin="%g" xmax="%g" ymin="%g" ymax="%g" zmin="%g" zmax="%g">'
    return grid_bcs(block3d(ncell=ncell, margin=margin, bc=False, **kw), [
        (box % (-1, n + 1, -1, n + 1, -1, lo + 0.01), 'dir="3" vel="0" id="-1"'),
        (box % (hi - 0.01, n + 1, -1, n + 1, -1, n + 1), 'dir="1" vel="0" id="-2"'),
        (box % (-1, n + 1, -1, n + 1, hi - 0.01, n + 1), 'dir="3" vel="-800" id="-3"'),
        (box % (-1, n + 1, -1, lo + 0.01, lo + 0.99, hi - 0.99), 'dir="12" angle="20" vel="0" id="-4"'),
    ])


def particle_bcs(xml, blocks):
    """Append a <ParticleBCs> element made of (shape opening tag, BC element) pairs, e.g.
    ('<BCBox xmin="0" ...>', '<TractionBC dir="11" face="6" style="1" stress="-5"/>')."""
    body = "".join("%s%s</%s>" % (shape, bc, shape[1:].split()[0].rstrip(">")) for shape, bc in blocks)
    return xml.replace("</MaterialPoints>", "</MaterialPoints><ParticleBCs>%s</ParticleBCs>" % body, 1)
