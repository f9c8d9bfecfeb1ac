"""``merge_crystal_maps()`` on the GPU.

Mirror of /root/reference/src/kikuchipy/indexing/_merge_crystal_maps.py:28-354: same keywords,
same ``ValueError`` texts, same phase-list bookkeeping; the array work (best phase per point,
values of the winning map, stable best-first ordering of every score of a point, unique
simulation indices) is one call into ``libkdi`` (``kdi_merge_crystal_maps``).

Maps are read by duck typing, so orix ``CrystalMap`` objects and this package's
:class:`~kikuchipy_b200.indexing.DictionaryIndexingResult` both work.  With orix installed the
result is an orix ``CrystalMap`` built exactly as at ``:349-364``; without it a
:class:`MergedCrystalMap` carrying the same arrays.

Differences from the reference, both on inputs where the reference itself misbehaves:

* one map point with ``mean_n_best > 1``: the reference's ``squeeze()`` (``:218``) also drops the
  point axis and the call fails later; here the mean over the first ``mean_n_best`` scores is used
  as for any other map size;
* ``abs(mean_n_best)`` larger than the number of scores per point is clipped to it (NumPy slicing
  does the same at ``:218``).
"""

from __future__ import annotations

import copy
import warnings
from math import copysign

import numpy as np

from . import _lib


class SimplePhase:
    """Stand-in for orix ``Phase`` when orix is absent: a name plus optional symmetry labels."""

    def __init__(self, name="", space_group=None, point_group=None):
        self.name = name
        self.space_group = space_group
        self.point_group = point_group

    def deepcopy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return f"SimplePhase(name={self.name!r}, space_group={self.space_group!r})"


class MergedCrystalMap:
    """What ``merge_crystal_maps`` returns when orix is not installed: ``phase_id``, ``rotations``
    (``(M, N, 4)`` or ``(M, 4)`` quaternions), ``prop`` (``scores``, ``merged_scores`` and, if
    asked for, the simulation-index arrays - also readable as attributes), ``phases``
    (``{id: phase}``), ``shape``, ``size``, ``scan_unit``, ``x`` / ``y``."""

    def __init__(self, rotations, phase_id, phases, prop, shape, step_sizes, scan_unit):
        self.rotations = rotations
        self.phase_id = phase_id
        self.phases = phases
        self.prop = prop
        self.shape = tuple(int(s) for s in shape)
        self.scan_unit = scan_unit
        self.is_in_data = np.ones(int(np.prod(self.shape)), dtype=bool)
        self.rotations_per_point = rotations.shape[1] if rotations.ndim == 3 else 1
        dx, dy = step_sizes
        if len(self.shape) == 1:
            self.x, self.y = np.arange(self.shape[0]) * dx, None
        else:
            ny, nx = self.shape
            self.x = np.tile(np.arange(nx) * dx, ny)
            self.y = np.repeat(np.arange(ny) * dy, nx)
        self.dx, self.dy = dx, dy

    @property
    def size(self):
        return self.phase_id.size

    def __getattr__(self, name):
        prop = self.__dict__.get("prop", {})
        if name in prop:
            return prop[name]
        raise AttributeError(name)

    def __repr__(self):
        names = ", ".join(f"{i}: {getattr(p, 'name', p)}" for i, p in self.phases.items())
        return f"MergedCrystalMap(shape={self.shape}, phases={{{names}}}, properties={list(self.prop)})"


# ---- duck-typed accessors ---------------------------------------------------------------------

def _is_in_data(xmap):
    return np.asarray(xmap.is_in_data, dtype=bool)


def _n_points(xmap):
    return int(_is_in_data(xmap).sum())


def _prop(xmap, name):
    a = np.asarray(xmap.prop[name])
    iid = _is_in_data(xmap)
    if a.shape[0] == iid.size and iid.size != iid.sum():  # full-size array held beside is_in_data
        a = a[iid]
    return a


def _rotation_data(xmap):
    if getattr(xmap, "rotations", None) is None:
        # an indexing result of a dictionary given without rotations: identity rotations, so that
        # scores and simulation indices can still be merged
        shape = _prop(xmap, next(iter(xmap.prop))).shape
        r = np.zeros(tuple(shape) + (4,))
        r[..., 0] = 1.0
        return r
    r = xmap.rotations
    r = np.asarray(r.data if hasattr(r, "data") and not isinstance(r, np.ndarray) else r, dtype=np.float64)
    iid = _is_in_data(xmap)
    if r.shape[0] == iid.size and iid.size != iid.sum():
        r = r[iid]
    return r


def _phase_id(xmap):
    pid = getattr(xmap, "phase_id", None)
    if pid is None:
        return np.zeros(_n_points(xmap), dtype=np.int64)
    pid = np.asarray(pid)
    iid = _is_in_data(xmap)
    if pid.shape[0] == iid.size and iid.size != iid.sum():
        pid = pid[iid]
    return pid


def _first_phase(xmap):
    """``xmap.phases_in_data[first id that is not -1].deepcopy()`` (``:243-247``)."""
    if hasattr(xmap, "phases_in_data"):
        phases = xmap.phases_in_data
        ids = list(phases.ids)
        if -1 in ids:
            ids.remove(-1)
        return phases[ids[0]].deepcopy()
    if hasattr(xmap, "phases") and isinstance(xmap.phases, dict):
        ids = [i for i in xmap.phases if i != -1]
        return copy.deepcopy(xmap.phases[ids[0]])
    return SimplePhase(getattr(xmap, "phase_name", "") or "")


def _equal_phase(p1, p2):
    """``signals/util/_crystal_map.py:65-108``, tolerant of phases without structure."""
    if p1.name != p2.name:
        return False, "names"
    sgs, pgs = [], []
    for p in (p1, p2):
        sg = getattr(p, "space_group", None)
        sgs.append(sg.number if hasattr(sg, "number") else (float(sg) if isinstance(sg, (int, float)) else np.nan))
        pg = getattr(p, "point_group", None)
        pgs.append(pg.data if hasattr(pg, "data") else np.nan)
    if not np.allclose(*sgs, equal_nan=True):
        return False, "space groups"
    if np.size(pgs[0]) != np.size(pgs[1]) or not np.allclose(*pgs, equal_nan=True):
        return False, "point groups"
    s1, s2 = getattr(p1, "structure", None), getattr(p2, "structure", None)
    if s1 is None or s2 is None:
        return True, None
    if len(s1) != len(s2):
        return False, "number of atoms"
    if not np.allclose(s1.lattice.abcABG(), s2.lattice.abcABG()):
        return False, "lattice parameters"
    for a1, a2 in zip(s1, s2):
        if a1.element != a2.element or not np.allclose(a1.xyz, a2.xyz) or not np.isclose(a1.occupancy, a2.occupancy):
            return False, "atoms"
    return True, None


class _PhaseBook:
    """The little of orix ``PhaseList`` the merge needs: sequential ids, lookup by name."""

    def __init__(self):
        self.phases = {}

    @property
    def names(self):
        return [p.name for p in self.phases.values()]

    def add_not_indexed(self):
        self.phases[-1] = SimplePhase("not_indexed")

    def add(self, phase):
        ids = [i for i in self.phases if i >= 0]
        self.phases[max(ids) + 1 if ids else 0] = phase

    def by_name(self, name):
        return next(p for p in self.phases.values() if p.name == name)

    def id_from_name(self, name):
        return next(i for i, p in self.phases.items() if p.name == name)


def merge_crystal_maps(
    crystal_maps,
    mean_n_best: int = 1,
    greater_is_better=None,
    scores_prop: str = "scores",
    simulation_indices_prop=None,
    navigation_masks=None,
    *,
    context=None,
):
    """Merge single-phase maps point by point on their scores (``_merge_crystal_maps.py:28``)."""
    n_maps = len(crystal_maps)

    # :96-105 masks from maps that do not hold every point
    if navigation_masks is None:
        if not all(_is_in_data(x).all() for x in crystal_maps):
            navigation_masks = []
            for x in crystal_maps:
                if hasattr(x, "_data_slices_from_coordinates"):
                    sl = x._data_slices_from_coordinates()
                    iid2d = _is_in_data(x).reshape(x._original_shape)[sl]
                else:
                    iid2d = _is_in_data(x).reshape(tuple(x.shape))
                navigation_masks.append(~iid2d)

    # :107-141 shapes
    if navigation_masks is not None:
        if len(navigation_masks) != n_maps:
            raise ValueError("Number of crystal maps and navigation masks must be equal")
        map_shapes = []
        for i, (mask, x) in enumerate(zip(navigation_masks, crystal_maps)):
            if isinstance(mask, np.ndarray):
                n_false, n_in = np.sum(~mask), _n_points(x)
                if n_false != n_in:
                    raise ValueError(
                        f"{i}. navigation mask does not have as many 'False', {n_false}, as there "
                        f"are points in the crystal map, {n_in}"
                    )
                map_shapes.append(mask.shape)
            elif mask is None:
                map_shapes.append(tuple(x.shape))
            else:
                raise ValueError(f"{i}. navigation mask must be a NumPy array or 'None'")
    else:
        map_shapes = [tuple(x.shape) for x in crystal_maps]
    if len({len(s) for s in map_shapes}) != 1 or not np.sum(abs(np.diff(map_shapes, axis=0))) == 0:
        raise ValueError("Crystal maps (and/or navigation masks) must have the same navigation shape")
    map_shape = tuple(int(s) for s in map_shapes[0])
    map_size = int(np.prod(map_shape))

    # :154-165 -> row of every map point inside each map (-1: the map does not hold the point)
    point_rows = [None] * n_maps
    if navigation_masks is not None:
        for i, mask in enumerate(navigation_masks):
            if mask is not None:
                keep = ~mask.ravel()
                rows = np.full(map_size, -1, dtype=np.int32)
                rows[keep] = np.arange(int(keep.sum()), dtype=np.int32)
                point_rows[i] = rows

    # :167-182
    per_point = [int(x.rotations_per_point) for x in crystal_maps]
    if not all(np.diff(per_point) == 0):
        raise ValueError("Crystal maps must have the same number of rotations and scores per point")
    n_scores = per_point[0]
    if simulation_indices_prop is not None:
        shp = np.shape(crystal_maps[0].prop[simulation_indices_prop])
        if len(shp) > 1 and shp[1] > n_scores:
            raise ValueError("Cannot merge maps with more simulation indices than scores per point")

    # :184-191
    if greater_is_better is None:
        sign = int(copysign(1, mean_n_best))
        mean_n_best = abs(mean_n_best)
    else:
        sign = 1 if greater_is_better else -1
    mean_n_best = max(1, min(int(mean_n_best), n_scores))

    scores = [_prop(x, scores_prop) for x in crystal_maps]
    scores_dtype = scores[0].dtype
    comp_dtype = np.dtype(f"f{scores_dtype.itemsize}") if scores_dtype.itemsize in (4, 8) else np.dtype(np.float64)
    scores2 = [np.asarray(s, dtype=comp_dtype).reshape(s.shape[0], n_scores) for s in scores]
    rots = [_rotation_data(x).reshape(s.shape[0], n_scores, 4) for x, s in zip(crystal_maps, scores2)]
    idx = None
    if simulation_indices_prop is not None:
        idx = [np.asarray(_prop(x, simulation_indices_prop)).reshape(s.shape[0], n_scores)
               for x, s in zip(crystal_maps, scores2)]
    # :227-237 (with masks the reference's assignment goes to a temporary: nothing is flagged)
    not_indexed = [None] * n_maps
    if navigation_masks is None:
        not_indexed = [(_phase_id(x) == -1) for x in crystal_maps]

    ctx = context if context is not None else _lib.default_context()
    out = ctx.merge_crystal_maps(scores2, rots, idx, point_rows, not_indexed, map_size, mean_n_best, sign,
                                 idx_as_double=navigation_masks is not None)
    phase_id = out["phase_id"]
    tail = (n_scores,) if n_scores > 1 else ()
    new_scores = out["scores"].reshape((map_size,) + tail).astype(scores_dtype, copy=False)
    new_rot = out["rotations"].reshape((map_size,) + tail + (4,))

    # :243-275 phase list of the merged map
    book = _PhaseBook()
    if -1 in phase_id:
        book.add_not_indexed()
    for i, x in enumerate(crystal_maps):
        phase_mask = phase_id == i
        if not phase_mask.any():
            continue
        phase = _first_phase(x)
        if phase.name in book.names:
            equal, different = _equal_phase(phase, book.by_name(phase.name))
            if equal:
                phase_id[phase_mask] = book.id_from_name(phase.name)
            else:
                name = phase.name
                phase.name = name + str(i)
                warnings.warn(
                    f"There are duplicates of phase '{name}' but the phases have different "
                    f"{different}, will therefore rename this phase's name to '{phase.name}' in "
                    "the merged PhaseList",
                )
                book.add(phase)
        else:
            book.add(phase)

    props = {scores_prop: new_scores, f"merged_{scores_prop}": out["merged_scores"]}
    if simulation_indices_prop is not None:
        props[simulation_indices_prop] = out["simulation_indices"].reshape((map_size,) + tail)
        props[f"merged_{simulation_indices_prop}"] = out["merged_simulation_indices"]

    first = crystal_maps[0]
    dx, dy = getattr(first, "dx", 1), getattr(first, "dy", 1)
    scan_unit = getattr(first, "scan_unit", "px")
    try:  # :349-364
        from orix.crystal_map import CrystalMap, PhaseList, create_coordinate_arrays
        from orix.quaternion import Rotation
    except ImportError:
        return MergedCrystalMap(new_rot, phase_id, book.phases, props, map_shape, (dx, dy), scan_unit)
    phase_list = PhaseList()
    for i, p in book.phases.items():
        if i == -1:
            phase_list.add_not_indexed()
        else:
            phase_list.add(p)
    coords, _ = create_coordinate_arrays(map_shape, step_sizes=(dx, dy)[: len(map_shape)])
    return CrystalMap(rotations=Rotation(new_rot), phase_id=phase_id, phase_list=phase_list, prop=props,
                      scan_unit=scan_unit, **coords)
