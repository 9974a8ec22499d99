"""Python front of the host model builder (include/vx3_model.h) + a VXA writer.

``ModelSpec`` holds what a VXA file holds (palette, lattice structure, environment, simulator options,
math-tree formulas).  ``build()`` feeds it to the native builder and returns the flat ``vx3_model_desc``;
``to_vxa()`` writes the same thing as VXA XML so the identical input can be given to the reference
(tests) or to the VXA reader.  No physics here — only marshalling.
"""
import ctypes as C

import numpy as np

from . import abi
from .libs import load_engine

MAT_DEFAULTS = dict(mat_model=0, elastic_mod=0.0, plastic_mod=0.0, yield_stress=0.0, fail_stress=0.0, fail_strain=0.0,
                    density=0.0, poissons_ratio=0.0, cte=0.0, material_temp_phase=0.0, u_static=0.0, u_dynamic=0.0,
                    is_pacemaker=0, is_measured=1, is_electrical_active=0, is_target=0, fixed=0, sticky=0,
                    pacemaker_period=0.0, signal_value_decay=0.9, signal_time_delay=0.03, inactive_period=0.03,
                    remove_after_s=0.0, thermal_on_after_s=0.0, cilia_on_after_s=0.0, cilia=0.0,
                    red=0.5, green=0.5, blue=0.5, alpha=1.0, name="Default",
                    n_data=0, strain_data=None, stress_data=None)  # MatModel 3 (MDL_DATA): <SSData> points, first one (0, 0)

# (VXA tag, spec key) for the <Mechanical> block, in the order the writer emits them
_MECH_TAGS = [("MatModel", "mat_model"), ("Elastic_Mod", "elastic_mod"), ("Plastic_Mod", "plastic_mod"),
              ("Yield_Stress", "yield_stress"), ("Fail_Stress", "fail_stress"), ("Fail_Strain", "fail_strain"),
              ("Density", "density"), ("Poissons_Ratio", "poissons_ratio"), ("CTE", "cte"),
              ("MaterialTempPhase", "material_temp_phase"), ("uStatic", "u_static"), ("uDynamic", "u_dynamic"),
              ("isPaceMaker", "is_pacemaker"), ("PaceMakerPeriod", "pacemaker_period"),
              ("signalValueDecay", "signal_value_decay"), ("signalTimeDelay", "signal_time_delay"),
              ("inactivePeriod", "inactive_period"), ("isMeasured", "is_measured"),
              ("RemoveFromSimulationAfterThisManySeconds", "remove_after_s"),
              ("TurnOnThermalExpansionAfterThisManySeconds", "thermal_on_after_s"),
              ("TurnOnCiliaAfterThisManySeconds", "cilia_on_after_s"), ("isElectricalActive", "is_electrical_active"),
              ("isTarget", "is_target"), ("Fixed", "fixed"), ("Sticky", "sticky"), ("Cilia", "cilia")]

ENV_DEFAULTS = dict(grav_enabled=1, grav_acc=-9.81, floor_enabled=1, temp_enabled=1, temp_base=25.0, temp_amplitude=0.0,
                    vary_temp_enabled=0, temp_period=0.1, bond_damping_z=0.0, col_damping_z=0.0, slow_damping_z=0.0,
                    volume_effects_enabled=0, self_col_enabled=0)

OPT_DEFAULTS = dict(dt_frac=0.9, enable_collision=1, enable_attach=0, enable_detach=0, watch_distance=1.0,
                    bounding_radius=0.75, safety_guard=500, record_step_size=0, record_link=0, record_voxel=1,
                    save_position_of_all_voxels=0, max_dist_in_voxel_lengths_to_count_as_pair=0.0, enable_cilia=0,
                    enable_signals=0, secondary_experiment=0, reinit_initial_position_after_s=0.0, enable_expansion=0)

PROG_TAGS = {abi.PROG_STOP: "StopCondition/StopConditionFormula", abi.PROG_FITNESS: "FitnessFunction",
             abi.PROG_FORCE_X: "ForceField/x_forcefield", abi.PROG_FORCE_Y: "ForceField/y_forcefield",
             abi.PROG_FORCE_Z: "ForceField/z_forcefield",
             abi.PROG_ATTACH_0: "AttachDetach/AttachCondition/Condition_0",
             abi.PROG_ATTACH_1: "AttachDetach/AttachCondition/Condition_1",
             abi.PROG_ATTACH_2: "AttachDetach/AttachCondition/Condition_2",
             abi.PROG_ATTACH_3: "AttachDetach/AttachCondition/Condition_3",
             abi.PROG_ATTACH_4: "AttachDetach/AttachCondition/Condition_4"}


def expr_to_tokens(expr):
    """Math-tree expression -> token list in the reference's order.

    ``expr`` is nested tuples, e.g. ``("SUB", ("VAR", "t"), ("CONST", 1))``.  The reference parser
    (ParseMathTree, src/VX3/VX3_SimulationManager.cu:157-275) walks the XML breadth-first, pushes each node on
    a stack under an mtEND sentinel and pops the stack: tokens = reverse(BFS order) + [END].
    """
    order, frontier = [], [expr]
    while frontier:
        nxt = []
        for node in frontier:
            order.append(node)
            nxt.extend(c for c in node[1:] if isinstance(c, tuple))
        frontier = nxt
    # the reference queue is one FIFO across levels: identical to level order above
    toks = []
    for node in reversed(order):
        op = node[0]
        if op == "VAR":
            toks.append((abi.OP["VAR"], float(abi.VARS[node[1]])))
        elif op == "CONST":
            toks.append((abi.OP["CONST"], float(node[1])))
        else:
            toks.append((abi.OP[op], 0.0))
    toks.append((abi.OP["END"], 0.0))
    return toks


def expr_to_xml(expr, indent="      "):
    op = expr[0]
    if op in ("VAR", "CONST"):
        return "%s<mt%s>%s</mt%s>\n" % (indent, op, repr(expr[1]) if op == "CONST" else expr[1], op)
    inner = "".join(expr_to_xml(c, indent + "  ") for c in expr[1:])
    return "%s<mt%s>\n%s%s</mt%s>\n" % (indent, op, inner, indent, op)


class ModelSpec:
    def __init__(self, lattice_dim=0.01, name="model"):
        self.lattice_dim = float(lattice_dim)
        self.name = name
        self.materials = []
        self.env = dict(ENV_DEFAULTS)
        self.opt = dict(OPT_DEFAULTS)
        self.has_damping = False  # whether <Damping> is written (absent => 0/0/0)
        self.structure = None     # uint8 [nz][ny][nx]
        self.phase_offset = None  # float64 [nz][ny][nx]
        self.base_cilia = None    # float64 [nz][ny][nx][3]
        self.shift_cilia = None
        self.programs = {}        # slot -> expr
        self.externals = []       # (voxel_index, dict)
        self._keep = []

    # ---- description ----
    def add_material(self, **kw):
        m = dict(MAT_DEFAULTS)
        for k in kw:
            if k not in m:
                raise KeyError(k)
        m.update(kw)
        self.materials.append(m)
        return len(self.materials)

    def set_env(self, **kw):
        for k in kw:
            if k not in self.env:
                raise KeyError(k)
        self.env.update(kw)
        if any(k in kw for k in ("bond_damping_z", "col_damping_z", "slow_damping_z")):
            self.has_damping = True

    def set_options(self, **kw):
        for k in kw:
            if k not in self.opt:
                raise KeyError(k)
        self.opt.update(kw)

    def set_structure(self, mat, phase_offset=None, base_cilia=None, shift_cilia=None):
        mat = np.ascontiguousarray(mat, dtype=np.uint8)
        assert mat.ndim == 3, "structure is [nz][ny][nx]"
        self.structure = mat
        self.phase_offset = None if phase_offset is None else np.ascontiguousarray(phase_offset, dtype=np.float64)
        self.base_cilia = None if base_cilia is None else np.ascontiguousarray(base_cilia, dtype=np.float64)
        self.shift_cilia = None if shift_cilia is None else np.ascontiguousarray(shift_cilia, dtype=np.float64)

    def set_program(self, slot, expr):
        self.programs[slot] = expr

    def set_external(self, voxel_index, dof_fixed=0, force=(0, 0, 0), moment=(0, 0, 0), translation=(0, 0, 0)):
        self.externals.append((voxel_index, dict(dof_fixed=dof_fixed, force=force, moment=moment, translation=translation)))

    # ---- native build ----
    def build(self, lib=None):
        """Returns (builder_handle, POINTER(ModelDesc)).  Keep the handle alive while the desc is used;
        free with lib.vx3_builder_destroy(handle)."""
        lib = lib or load_engine()
        b = lib.vx3_builder_create(self.lattice_dim)
        lib.vx3_builder_set_name(b, self.name.encode())
        for m in self.materials:
            p = abi.MaterialParams()
            lib.vx3_material_params_default(C.byref(p))
            for k, v in m.items():
                if k in ("name", "strain_data", "stress_data", "n_data"):
                    continue
                setattr(p, k, v)
            if m["strain_data"] is not None:
                n = len(m["strain_data"])
                sd, ss = (C.c_double * n)(*m["strain_data"]), (C.c_double * n)(*m["stress_data"])
                self._keep += [sd, ss]
                p.n_data, p.strain_data, p.stress_data = n, sd, ss
            lib.vx3_builder_add_material(b, C.byref(p))
        e = abi.EnvParams()
        lib.vx3_env_params_default(C.byref(e))
        for k, v in self.env.items():
            setattr(e, k, v)
        lib.vx3_builder_set_env(b, C.byref(e))
        o = abi.SimOptions()
        lib.vx3_sim_options_default(C.byref(o))
        for k, v in self.opt.items():
            setattr(o, k, v)
        lib.vx3_builder_set_options(b, C.byref(o))
        for slot, expr in self.programs.items():
            toks = expr_to_tokens(expr)
            arr = (abi.Token * len(toks))()
            for i, (op, val) in enumerate(toks):
                arr[i].op, arr[i].value = op, val
            lib.vx3_builder_set_program(b, slot, arr, len(toks))
        nz, ny, nx = self.structure.shape
        dp = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        rc = lib.vx3_builder_set_structure(b, nx, ny, nz, self.structure.ctypes.data_as(C.POINTER(C.c_uint8)),
                                           dp(self.phase_offset), dp(self.base_cilia), dp(self.shift_cilia))
        if rc != 0:
            raise ValueError("vx3_builder_set_structure failed")
        for idx, ext in self.externals:
            x = abi.External()
            x.dof_fixed = ext["dof_fixed"]
            for k in range(3):
                x.force[k] = ext["force"][k]
                x.moment[k] = ext["moment"][k]
                x.translation[k] = ext["translation"][k]
            x.rotation_q[0] = 1.0
            lib.vx3_builder_set_external(b, idx, C.byref(x))
        d = lib.vx3_builder_build(b)
        if not d:
            msg = lib.vx3_model_last_error().decode()
            lib.vx3_builder_destroy(b)
            raise ValueError("vx3_builder_build failed: " + msg)
        return b, d

    # ---- VXA text (same schema as the reference's demos/basic/base.vxa) ----
    def to_vxa(self):
        g = lambda v: repr(float(v)) if isinstance(v, float) else str(int(v))
        o, e = self.opt, self.env
        s = ['<VXA Version="1.1">\n<Simulator>\n']
        s.append("  <Integration><DtFrac>%s</DtFrac></Integration>\n" % g(float(o["dt_frac"])))
        if self.has_damping:
            s.append("  <Damping><BondDampingZ>%s</BondDampingZ><ColDampingZ>%s</ColDampingZ><SlowDampingZ>%s</SlowDampingZ></Damping>\n"
                     % (g(float(e["bond_damping_z"])), g(float(e["col_damping_z"])), g(float(e["slow_damping_z"]))))
        if e["volume_effects_enabled"]:
            s.append("  <Features><VolumeEffectsEnabled>1</VolumeEffectsEnabled></Features>\n")
        s.append("  <RecordHistory><RecordStepSize>%d</RecordStepSize><RecordVoxel>%d</RecordVoxel><RecordLink>%d</RecordLink></RecordHistory>\n"
                 % (o["record_step_size"], o["record_voxel"], o["record_link"]))
        tree = {}
        for slot, expr in self.programs.items():
            node = tree
            parts = PROG_TAGS[slot].split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = expr
        ad = tree.setdefault("AttachDetach", {})
        ad.update({"EnableCollision": o["enable_collision"], "EnableAttach": o["enable_attach"], "EnableDetach": o["enable_detach"],
                   "watchDistance": float(o["watch_distance"]), "boundingRadius": float(o["bounding_radius"]),
                   "SafetyGuard": o["safety_guard"]})

        def emit(node, ind):
            out = ""
            for k, v in node.items():
                if isinstance(v, dict):
                    out += "%s<%s>\n%s%s</%s>\n" % (ind, k, emit(v, ind + "  "), ind, k)
                elif isinstance(v, tuple):
                    out += "%s<%s>\n%s%s</%s>\n" % (ind, k, expr_to_xml(v, ind + "  "), ind, k)
                else:
                    out += "%s<%s>%s</%s>\n" % (ind, k, g(v), k)
            return out
        s.append(emit(tree, "  "))
        for tag, key in (("SavePositionOfAllVoxels", "save_position_of_all_voxels"),
                         ("MaxDistInVoxelLengthsToCountAsPair", "max_dist_in_voxel_lengths_to_count_as_pair"),
                         ("EnableCilia", "enable_cilia"), ("EnableSignals", "enable_signals"),
                         ("SecondaryExperiment", "secondary_experiment"),
                         ("ReinitializeInitialPositionAfterThisManySeconds", "reinit_initial_position_after_s"),
                         ("EnableExpansion", "enable_expansion")):
            s.append("  <%s>%s</%s>\n" % (tag, g(o[key]), tag))
        s.append("</Simulator>\n<Environment>\n")
        s.append("  <Gravity><GravEnabled>%d</GravEnabled><GravAcc>%s</GravAcc><FloorEnabled>%d</FloorEnabled></Gravity>\n"
                 % (e["grav_enabled"], g(float(e["grav_acc"])), e["floor_enabled"]))
        s.append("  <Thermal><TempEnabled>%d</TempEnabled><TempAmplitude>%s</TempAmplitude><TempBase>%s</TempBase>"
                 "<VaryTempEnabled>%d</VaryTempEnabled><TempPeriod>%s</TempPeriod></Thermal>\n"
                 % (e["temp_enabled"], g(float(e["temp_amplitude"])), g(float(e["temp_base"])), e["vary_temp_enabled"],
                    g(float(e["temp_period"]))))
        s.append("</Environment>\n<VXC Version=\"0.94\">\n  <Lattice><Lattice_Dim>%s</Lattice_Dim></Lattice>\n  <Palette>\n"
                 % g(self.lattice_dim))
        for i, m in enumerate(self.materials):
            s.append('    <Material ID="%d">\n      <Name>%s</Name>\n' % (i + 1, m["name"]))
            s.append("      <Display><Red>%s</Red><Green>%s</Green><Blue>%s</Blue><Alpha>%s</Alpha></Display>\n"
                     % tuple(g(float(m[k])) for k in ("red", "green", "blue", "alpha")))
            s.append("      <Mechanical>\n")
            if m["strain_data"] is not None:
                s.append("        <SSData><NumDataPts>%d</NumDataPts><StrainData>%s</StrainData><StressData>%s</StressData></SSData>\n"
                         % (len(m["strain_data"]), "".join("<Strain>%s</Strain>" % g(float(v)) for v in m["strain_data"]),
                            "".join("<Stress>%s</Stress>" % g(float(v)) for v in m["stress_data"])))
            for tag, key in _MECH_TAGS:
                s.append("        <%s>%s</%s>\n" % (tag, g(m[key]), tag))
            s.append("      </Mechanical>\n    </Material>\n")
        nz, ny, nx = self.structure.shape
        s.append('  </Palette>\n  <Structure Compression="ASCII_READABLE">\n    <X_Voxels>%d</X_Voxels><Y_Voxels>%d</Y_Voxels><Z_Voxels>%d</Z_Voxels>\n    <Data>\n'
                 % (nx, ny, nz))
        for z in range(nz):
            s.append("      <Layer><![CDATA[%s]]></Layer>\n" % "".join(chr(48 + int(v)) for v in self.structure[z].ravel()))
        s.append("    </Data>\n")
        if self.phase_offset is not None:
            s.append("    <PhaseOffset>\n")
            for z in range(nz):
                s.append("      <Layer><![CDATA[%s]]></Layer>\n" % ",".join(repr(float(v)) for v in self.phase_offset[z].ravel()))
            s.append("    </PhaseOffset>\n")
        for tag, arr in (("BaseCiliaForce", self.base_cilia), ("ShiftCiliaForce", self.shift_cilia)):
            if arr is not None:
                s.append("    <%s>\n" % tag)
                for z in range(nz):
                    s.append("      <Layer><![CDATA[%s]]></Layer>\n" % ",".join(repr(float(v)) for v in arr[z].ravel()))
                s.append("    </%s>\n" % tag)
        s.append("  </Structure>\n</VXC>\n</VXA>\n")
        return "".join(s)


# ---------------------------------------------------------------- state helpers (host buffers for vx3_state_view)
class StateBuffers:
    """numpy-backed buffers for a vx3_state_view."""
    F3 = ["pos", "lin_mom", "ang_mom", "contact_force"]
    L3 = ["link_pos2", "link_angle1v", "link_angle2v", "link_force_neg", "link_force_pos", "link_moment_neg", "link_moment_pos"]

    def __init__(self, n_voxels, n_links):
        nv, nl = max(n_voxels, 1), max(n_links, 1)
        self.cap = (n_voxels, n_links)
        self.a = {}
        for k in self.F3:
            self.a[k] = np.zeros((nv, 3))
        self.a["orient"] = np.zeros((nv, 4))
        self.a["vox_flags"] = np.zeros(nv, np.int32)
        self.a["temp"] = np.zeros(nv, np.float32)
        self.a["vox_links"] = np.zeros((nv, 6), np.int32)
        for k in ("link_vneg", "link_vpos", "link_axis", "link_mat", "link_flags"):
            self.a[k] = np.zeros(nl, np.int32)
        for k in self.L3:
            self.a[k] = np.zeros((nl, 3))
        for k in ("link_strain", "link_max_strain", "link_strain_offset", "link_stress"):
            self.a[k] = np.zeros(nl, np.float32)
        self.a["link_rest_length"] = np.zeros(nl)
        self.a["signal"] = np.zeros((nv, 6))
        self.view = abi.StateView()
        self.view.n_voxels, self.view.n_links = n_voxels, n_links
        for name, ctype in abi.StateView._fields_:
            if name in self.a:
                setattr(self.view, name, self.a[name].ctypes.data_as(ctype))

    def result(self):
        nv, nl = self.view.n_voxels, self.view.n_links
        out = {}
        for k, v in self.a.items():
            out[k] = v[:nl].copy() if k.startswith("link_") else v[:nv].copy()
        return out
