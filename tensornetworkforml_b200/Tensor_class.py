"""Named-axis tensor: host-side mirror of the reference's ``Tensor`` (TC:6-298).

Same constructor, attributes (``elem``, ``shape``, ``rank``, ``axes_names``, ``aggregations``,
``history_axes_names``) and methods, so objects pickled by the reference load into this class and objects
pickled here load into the reference's.  It is bookkeeping only (reshape / transpose / names); the arithmetic
of the hot path lives in libtnml.so.
"""
from __future__ import annotations

import numpy as np


class Tensor:
    def __init__(self, elem=None, shape=None, axes_names=None, scale=1.):
        if elem is not None:
            self.elem = elem
        elif shape is not None:
            # same RNG call as TC:63-64 so seeded constructions reproduce the reference's weights
            self.elem = np.random.random(size=shape)
            self.elem /= scale
        else:
            raise Exception('You have to provide either the elements of the tensor or its shape')
        self.shape = self.elem.shape
        self.rank = len(self.shape)
        self.aggregations = {}
        self.axes_names = None
        if axes_names is not None:
            try:
                n_names = len(axes_names)
            except TypeError:
                print("=== Warning ===\nThe object that describes the indexes names have at least to support the "
                      "built-in len function.\naxes_names attribute has not been inizialized.")
                return
            if n_names != self.rank:
                print("=== Warning ===\nThe number of names should match the rank of the tensor."
                      "\naxes_names attribute has not been inizialized.")
                return
            self.axes_names = np.array(axes_names)
            self.history_axes_names = [np.array(axes_names)]

    # ---- bookkeeping -----------------------------------------------------------------------------
    def update_members(self, axes_names):
        """Refresh names / shape / rank after ``elem`` changed (TC:244-256)."""
        self.axes_names = np.array(axes_names)
        self.shape = self.elem.shape
        self.rank = len(self.shape)

    def ax_to_index(self, axes):
        """Position(s) of the named axis / axes (TC:219-241)."""
        if type(axes) == str:
            return np.where(self.axes_names == axes)[0][0]
        return [np.where(self.axes_names == name)[0][0] for name in axes]

    def transpose(self, permutation):
        """Reorder the axes to the given order of names (TC:202-216)."""
        order = self.ax_to_index(permutation)
        self.elem = np.transpose(self.elem, order)
        self.update_members(permutation)

    # ---- aggregate / disaggregate (TC:97-199) --------------------------------------------------------
    def aggregate(self, axes_names=None, new_ax_name=None, debug=False):
        """Fuse the listed axes (in the listed order, row-major) into one leading axis ``new_ax_name``."""
        if new_ax_name is None:
            raise ValueError("You have to provide the name of the new axes")
        if self.axes_names is None:
            raise ValueError("This function can be called only if the axes names are defined")
        if axes_names is None:
            axes_names = self.axes_names
        for name in axes_names:
            assert name in self.axes_names, "The " + name + " axes wasn't found in the tensor"
        fused = self.ax_to_index(axes_names)
        kept = [i for i in range(self.rank) if i not in set(fused)]
        if debug:
            print("Aggregating...", fused, kept)
        self.aggregations[new_ax_name] = dict(zip(axes_names, np.array(self.shape)[fused]))
        kept_sizes = [self.shape[i] for i in kept]
        kept_names = self.axes_names[kept]
        self.elem = np.transpose(self.elem, fused + kept).reshape([-1] + kept_sizes)
        self.update_members(np.concatenate([[new_ax_name], kept_names]))

    def disaggregate(self, ax):
        """Undo ``aggregate``: the fused axis is moved to the front and split back into its members."""
        assert ax in self.axes_names, "The " + ax + " ax wasn't found in the tensor."
        assert ax in self.aggregations.keys(), "The " + ax + " does not represent an aggregated ax."
        members = self.aggregations[ax]
        pos = self.ax_to_index(ax)
        order = [pos] + [i for i in range(self.rank) if i != pos]
        self.elem = np.transpose(self.elem, order)
        self.update_members(self.axes_names[order])
        self.elem = self.elem.reshape(list(members.values()) + list(self.shape[1:]))
        self.update_members(np.concatenate([list(members.keys()), self.axes_names[1:]]))
        self.aggregations.pop(ax)

    # ---- misc ------------------------------------------------------------------------------------------
    def check_names(self):
        print("=" * 10 + "axes_names type" + "=" * 10)
        print(type(self.axes_names))

    def __str__(self):
        print("=" * 10 + " Tensor description " + "=" * 10)
        print("Tensor shape: ", self.shape)
        print("Tensor rank: ", self.rank)
        print("Axes names: ", self.axes_names)
        return ""

    def _aligned(self, other):
        assert np.all(np.isin(self.axes_names, other.axes_names)), "Error: axes don't match, cannot sum tensors."
        other.transpose(self.axes_names)          # the reference also permutes the right operand in place (TC:282)
        return other.elem

    def __add__(self, o):
        return Tensor(elem=self.elem + self._aligned(o), axes_names=self.axes_names)

    def __sub__(self, o):
        return Tensor(elem=self.elem - self._aligned(o), axes_names=self.axes_names)


Tensor.__module__ = "Tensor_class"      # pickles stay interchangeable with the reference
