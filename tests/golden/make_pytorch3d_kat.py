"""Known-answer vectors for the three pytorch3d 0.7.2 functions on the path, derived from the PUBLISHED FORMULAS in
plain float64 `math` -- independently of oracle/pytorch3d_restated.py (no torch, no shared code).

pytorch3d itself is not in this image and not vendored by the reference (README.md:48 pins 0.7.2), so these are
NOT outputs of pytorch3d: row a16 stays "parity unpinned".  What the vectors pin is that the oracle restatement
and the K4 kernel implement the documented algorithm, at the places where it is easy to get wrong:

  so3_relative_angle(R1, R2) = acos_linear_extrapolation((tr(R1 R2^T) - 1) / 2, bounds = +-(1 - 1e-4)),
      ValueError when tr < -1 - 1e-4 or tr > 3 + 1e-4
  acos_linear_extrapolation(x, (lo, hi)) = acos(x) for lo < x < hi, else the first-order Taylor polynomial of acos
      about the violated bound b:  acos(b) + (x - b) * (-1 / sqrt(1 - b^2))
      => identical rotations (x = 1): acos(1-1e-4) - 1e-4 / sqrt(1 - (1-1e-4)^2) = 0.0070711 rad = 0.40514 deg
  matrix_to_quaternion(M): q_abs = sqrt(max(0, 1 +- m00 +- m11 +- m22)); four candidate quaternions, each divided by
      2 * max(q_abs_i, 0.1); the candidate with the largest q_abs is returned (real part first, no sign fix-up)
  quaternion_to_matrix(q): two_s = 2 / |q|^2 (so non-unit quaternions give the rotation of their direction)

usage: python tests/golden/make_pytorch3d_kat.py   (writes tests/golden/pytorch3d_kat.json)
"""
import json
import math
import os

B = 1.0 - 1e-4


def acos_ext(x):
    if -B < x < B:
        return math.acos(x)
    b = B if x >= B else -B
    return math.acos(b) + (x - b) * (-1.0 / math.sqrt(1.0 - b * b))


def rot(axis, t):
    c, s = math.cos(t), math.sin(t)
    if axis == "x":
        return [[1, 0, 0], [0, c, -s], [0, s, c]]
    if axis == "y":
        return [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    return [[c, -s, 0], [s, c, 0], [0, 0, 1]]


def main():
    kat = {"note": "derived from the published pytorch3d 0.7.2 formulas in float64; NOT pytorch3d outputs (parity unpinned)"}
    kat["acos_linear_extrapolation"] = [
        {"x": x, "y": acos_ext(x)} for x in (1.0, B, 1.0 - 5e-5, 1.0 + 5e-5, 0.5, 0.0, -0.5, -B, -1.0, -1.0 - 5e-5, B - 1e-9, -B + 1e-9)]
    angles = [0.0, 1e-3, 0.01, 0.0141, 0.0142, 0.5, math.pi / 2, 3.0, math.pi - 0.0142, math.pi - 0.0141, math.pi - 0.01, math.pi]
    # R1 = Rz(theta), R2 = I  =>  tr(R1 R2^T) = 1 + 2 cos(theta)
    kat["so3_relative_angle_about_z"] = [{"theta": t, "R1": rot("z", t), "angle": acos_ext(((1 + 2 * math.cos(t)) - 1) / 2)} for t in angles]
    kat["identity_angle_rad"] = acos_ext(1.0)
    kat["identity_angle_deg"] = math.degrees(acos_ext(1.0))
    kat["trace_out_of_range"] = [[[1.0004, 0, 0], [0, 1.0004, 0], [0, 0, 1.0004]],      # tr = 3.0012 > 3 + 1e-4
                                 [[-1, 0, 0], [0, -1, 0], [0, 0, 0.9998]]]              # tr = -1.0002 < -1 - 1e-4
    kat["trace_in_range_edge"] = [[[1.00002, 0, 0], [0, 1.00002, 0], [0, 0, 1.00002]]]  # tr = 3.00006: allowed, extrapolated
    # matrix_to_quaternion: one rotation per branch (largest component w, x, y, z)
    h = 0.1
    kat["matrix_to_quaternion"] = [
        {"M": rot("z", 0.3), "q": [math.cos(0.15), 0.0, 0.0, math.sin(0.15)], "branch": "w"},
        {"M": rot("x", math.pi - 2 * h), "q": [math.sin(h), math.cos(h), 0.0, 0.0], "branch": "x"},
        {"M": rot("y", math.pi - 2 * h), "q": [math.sin(h), 0.0, math.cos(h), 0.0], "branch": "y"},
        {"M": rot("z", math.pi - 2 * h), "q": [math.sin(h), 0.0, 0.0, math.cos(h)], "branch": "z"},
        # not rotations: all four radicands equal (argmax takes the first) / the 0.1 floor on a zero candidate
        {"M": [[0, 0, 0], [0, 0, 0], [0, 0, 0]], "q": [0.5, 0.0, 0.0, 0.0], "branch": "w (tie)"},
        {"M": [[0.01, 0, 0], [0, 0.01, 0], [0, 0, 0.01]], "q": [1.03 / (2 * math.sqrt(1.03)), 0.0, 0.0, 0.0], "branch": "w"},
    ]
    s = 1 / math.sqrt(2)
    kat["quaternion_to_matrix"] = [
        {"q": [2.0, 0.0, 0.0, 0.0], "M": [[1, 0, 0], [0, 1, 0], [0, 0, 1]]},
        {"q": [3 * s, 3 * s, 0.0, 0.0], "M": rot("x", math.pi / 2)},
        {"q": [math.cos(0.4), 0.0, math.sin(0.4), 0.0], "M": rot("y", 0.8)},
    ]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pytorch3d_kat.json")
    with open(path, "w") as f:
        json.dump(kat, f, indent=1)
    print(path, kat["identity_angle_rad"], kat["identity_angle_deg"])


if __name__ == "__main__":
    main()
