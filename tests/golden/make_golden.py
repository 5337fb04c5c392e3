"""Generate tests/golden/*.npz from the REFERENCE package itself.

Run in the build container only (the GPU box has no /root/reference):

    cp -r /root/reference/{elasticdeform,setup.py,README.md} /tmp/refbuild
    (cd /tmp/refbuild && python setup.py build_ext --inplace)
    PYTHONPATH=/tmp/refbuild python tests/golden/make_golden.py

Every case is produced by the unmodified reference package
(elasticdeform.deform_grid / deform_grid_gradient, reference deform_grid.py:52, :182)
on seeded inputs; the .npz stores inputs, the call's keyword arguments (as JSON) and
the reference outputs.  The reference ships no golden vectors or seeds of its own
(its tests draw unseeded random data), so these fixtures are what pins the oracle
and the CUDA path to the reference's numbers.  Cases are small twins of the five
BASELINE.json configs plus the mode/order/dtype matrix of the reference's tests.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _crop_to_json(crop):
    return None if crop is None else [[s.start, s.stop] for s in crop]


def cases():
    rng = np.random.default_rng(20240917)
    out = []

    def add(name, X, D, grad=True, **kw):
        out.append((name, X, D, grad, kw))

    # cfg1 twin: 2-D float32, 3x3 grid, order 3, reflect (full size: it is tiny)
    add("cfg1_2d_200x300_f32_o3_reflect", rng.random((200, 300), dtype=np.float32),
        rng.standard_normal((2, 3, 3)) * 25, order=3, mode='reflect')
    # cfg2 twin: 3-D float32, 5^3 grid, order 3 (32^3 instead of 128^3)
    add("cfg2_3d_32_f32_o3", rng.random((32, 32, 32), dtype=np.float32),
        rng.standard_normal((3, 5, 5, 5)) * 3, order=3)
    add("cfg2_3d_32_f32_o3_noprefilter", rng.random((32, 32, 32), dtype=np.float32),
        rng.standard_normal((3, 5, 5, 5)) * 3, order=3, prefilter=False)
    # cfg3 twin: image + int32 label pair, orders [3, 0]
    add("cfg3_3d_pair_f32_i32", [rng.random((24, 28, 32), dtype=np.float32),
                                 rng.integers(0, 5, (24, 28, 32), dtype=np.int32)],
        rng.standard_normal((3, 5, 5, 5)) * 3, order=[3, 0])
    # cfg4 twin: crop + 3-D affine (rotation about axis 0 by 15 deg, zoom 1.2, about the crop centre)
    th = np.radians(15.0)
    R = np.array([[1, 0, 0], [0, np.cos(th), -np.sin(th)], [0, np.sin(th), np.cos(th)]]) * 1.2
    c = np.array([9.5, 9.5, 9.5])
    A = np.concatenate([R, (c - R @ c)[:, None]], axis=1)
    add("cfg4_3d_crop_affine", rng.random((40, 40, 40), dtype=np.float32),
        rng.standard_normal((3, 5, 5, 5)) * 2, order=3,
        crop=(slice(10, 30), slice(10, 30), slice(10, 30)), affine=A)
    # cfg5 twin: multi-channel 4-D, axis=(1,2,3), order 1
    add("cfg5_4d_channels_o1", rng.random((3, 20, 22, 24), dtype=np.float32),
        rng.standard_normal((3, 5, 5, 5)) * 2, order=1, axis=(1, 2, 3))
    # mode x order matrix, 2-D float64 (reference test_basic_2d / test_grad_2d)
    for mode in ('nearest', 'wrap', 'reflect', 'mirror', 'constant'):
        for order in (0, 1, 2, 3, 4, 5):
            add("modes_2d_f64_%s_o%d" % (mode, order), rng.random((30, 25)),
                rng.standard_normal((2, 3, 5)) * 6, order=order, mode=mode, cval=0.5)
    # float32 at every order, 3-D
    for order in (0, 1, 2, 3, 4, 5):
        add("f32_3d_o%d" % order, rng.random((18, 20, 22), dtype=np.float32),
            rng.standard_normal((3, 3, 4, 5)) * 3, order=order, mode='mirror')
    # dtypes (rounding / clamping rules of deform.c:292-306)
    for dt in ('uint8', 'int16', 'int32', 'int64', 'uint16', 'float64'):
        X = (rng.random((26, 31)) * 300 - 20).astype(dt)
        add("dtype_%s_o1" % dt, X, rng.standard_normal((2, 3, 3)) * 5, order=1, cval=7.6)
        add("dtype_%s_o0" % dt, X, rng.standard_normal((2, 3, 3)) * 5, order=0, cval=-3.4)
    # rotate / zoom / affine / crop, 2-D (reference test_crop_rotate_zoom)
    add("rotzoom_2d", rng.random((60, 50)), rng.standard_normal((2, 3, 3)) * 3,
        rotate=30, zoom=1.5, crop=(slice(10, 50), slice(5, 45)), affine=np.eye(3))
    # multi input, different axes, strided (F-order) input, list-valued mode / cval
    add("multi_axes", [np.asfortranarray(rng.random((3, 20, 30))), rng.random((20, 30)).astype(np.float32)],
        rng.standard_normal((2, 5, 3)) * 4, order=[2, 3], mode=['constant', 'reflect'],
        cval=[0.0, 1.0], axis=[(1, 2), (0, 1)])
    # degenerate control grid (1 point on an axis) and 1-D
    add("grid_1x5", rng.random((31, 17)), rng.standard_normal((2, 1, 5)) * 4, order=3)
    add("one_d", rng.random((50,)), rng.standard_normal((1, 4)) * 4, order=3, mode='wrap')
    # zero displacement (coordinates exactly on the integer lattice)
    add("identity_3d_f32", rng.random((12, 14, 16), dtype=np.float32), np.zeros((3, 3, 3, 3)), order=1)
    return out


def main():
    import elasticdeform                                   # the reference package
    assert "refbuild" in elasticdeform.__file__ or "reference" in elasticdeform.__file__, elasticdeform.__file__
    rng = np.random.default_rng(7)
    index = []
    for name, X, D, grad, kw in cases():
        Y = elasticdeform.deform_grid(X, D, **kw)
        Ys = Y if isinstance(Y, list) else [Y]
        Xs = X if isinstance(X, list) else [X]
        data = {"displacement": D}
        for i, (x, y) in enumerate(zip(Xs, Ys)):
            data["x%d" % i] = x
            data["y%d" % i] = y
        if grad:
            dYs = [(rng.random(y.shape) * 4).astype(y.dtype) for y in Ys]
            gkw = dict(kw)
            gkw["X_shape"] = [x.shape for x in Xs] if isinstance(X, list) else Xs[0].shape
            dX = elasticdeform.deform_grid_gradient(dYs if isinstance(X, list) else dYs[0], D, **gkw)
            dXs = dX if isinstance(dX, list) else [dX]
            for i, (dy, dx) in enumerate(zip(dYs, dXs)):
                data["dy%d" % i] = dy
                data["dx%d" % i] = dx
        meta = dict(kw)
        if "crop" in meta:
            meta["crop"] = _crop_to_json(meta["crop"])
        if "affine" in meta:
            meta["affine"] = np.asarray(meta["affine"]).tolist()
        if "axis" in meta and isinstance(meta["axis"], list):
            meta["axis"] = [list(a) for a in meta["axis"]]
        elif "axis" in meta and isinstance(meta["axis"], tuple):
            meta["axis"] = {"tuple": list(meta["axis"])}
        data["meta"] = np.array(json.dumps({"kwargs": meta, "is_list": isinstance(X, list),
                                            "n": len(Xs), "grad": grad}))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        index.append(name)
    with open(os.path.join(HERE, "INDEX.txt"), "w") as f:
        f.write("\n".join(index) + "\n")
    print("wrote %d fixtures" % len(index))


if __name__ == "__main__":
    main()
