"""config.xml / <CELL>.xml / <CELL>.pos writers for the facade tests (values = the reference's material files,
SURVEY.md section 8d: identical in every named config)"""
RBC_XML = """<?xml version="1.0" ?>
<hemocell><MaterialModel>
  <name>RBC</name><eta_m> 0.0 </eta_m><kBend> 80.0 </kBend><kVolume> 20.0 </kVolume><kArea> 5.0 </kArea><kLink> 15.0 </kLink>
  <minNumTriangles> 600 </minNumTriangles><radius> 3.91e-6 </radius><Volume> 90 </Volume>
</MaterialModel></hemocell>
"""


def shear_config(tmax, tmeas, material_every=1, particle_every=1, height_um=10.0, shearrate=111.0, dt=0.5e-7, warmup=0):
    return f"""<?xml version="1.0" ?>
<hemocell>
<parameters><warmup> {warmup} </warmup><outputDirectory>tmp</outputDirectory></parameters>
<ibm><radius> 3.91e-6 </radius><cellType>RBC</cellType><stepMaterialEvery> {material_every} </stepMaterialEvery>
     <stepParticleEvery> {particle_every} </stepParticleEvery></ibm>
<domain><height> {height_um} </height><shearrate> {shearrate} </shearrate><rhoP> 1025 </rhoP><nuP> 1.1e-6 </nuP><dx> 0.5e-6 </dx>
        <dt> {dt} </dt><particleEnvelope>20</particleEnvelope><kBT>4.100531391e-21</kBT></domain>
<sim><tmax> {tmax} </tmax><tmeas> {tmeas} </tmeas><tcheckpoint> 100000000 </tcheckpoint></sim>
</hemocell>
"""


def write_shear_case(d, tmax, tmeas, rows=((9.5, 9.5, 4.5, 90, 0, 0),), **kw):
    (d / "config.xml").write_text(shear_config(tmax, tmeas, **kw))
    (d / "RBC.xml").write_text(RBC_XML)
    (d / "RBC.pos").write_text(f"{len(rows)}\n" + "".join(" ".join(str(v) for v in r) + "\n" for r in rows))
