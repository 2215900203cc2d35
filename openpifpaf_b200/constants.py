"""Dataset topology constants consumed by the decoder as *configuration*.

The decoder itself never embeds a skeleton: it receives `skeleton` (1-based
pairs, like ``headmeta.Caf.skeleton``) from the head metas exactly as the
reference does (decoder/cifcaf.py:119-122 passes ``skeleton - 1`` to C++).
These tables exist for the synthetic workloads of tests/ and bench.py and are
checked for equality against the reference plugin constants
(plugins/coco/constants.py:4-8, plugins/wholebody/constants.py:4-38) by
tests/test_constants.py when /root/reference is present.
"""

# COCO person keypoint topology, 17 keypoints / 19 associations (1-based).
COCO_PERSON_SKELETON = [
    (16, 14), (14, 12), (17, 15), (15, 13), (12, 13), (6, 12), (7, 13),
    (6, 7), (6, 8), (7, 9), (8, 10), (9, 11), (2, 3), (1, 2), (1, 3),
    (2, 4), (3, 5), (4, 6), (5, 7),
]
COCO_N_KEYPOINTS = 17


def _chain(first, last):
    """(first, first+1), ..., (last-1, last)"""
    return [(x, x + 1) for x in range(first, last)]


def _hand(root, wrist):
    starts = [root + 1 + 4 * k for k in range(5)]          # five finger roots
    out = [(root, wrist)] + [(root, s) for s in starts]
    for s in starts:
        out += _chain(s, s + 3)
    out += [(starts[0] + 1, starts[1])] + [(a, b) for a, b in zip(starts[1:], starts[2:])]
    return out


def wholebody_skeleton():
    """COCO-WholeBody topology: 133 keypoints (17 body, 6 feet, 68 face, 2x21 hands),
    160 associations, 1-based; same edge order as the reference plugin."""
    body_foot = COCO_PERSON_SKELETON + [(16, 20), (16, 19), (16, 18), (17, 23), (17, 21), (17, 22)]
    face = (
        [(25, 5), (39, 4), (54, 1), (60, 3), (3, 63), (66, 2), (2, 69)]
        + _chain(24, 40)
        + [(24, 41)] + _chain(41, 45) + [(45, 51)]
        + [(40, 50), (50, 49), (49, 48), (48, 47), (47, 46), (46, 51)]
        + [(24, 60)] + _chain(60, 63) + [(63, 51), (63, 64), (64, 65), (65, 60)]
        + [(40, 69), (69, 68), (68, 67), (67, 66), (66, 51), (66, 71), (71, 70), (70, 69)]
        + _chain(51, 59)
        + [(59, 54), (57, 75), (78, 36), (72, 28), (72, 83)] + _chain(72, 83)
        + [(72, 84)] + _chain(84, 88) + [(88, 78)]
        + [(72, 91), (91, 90), (90, 89), (89, 78)]
    )
    return body_foot + face + _hand(92, 10) + _hand(113, 11)


WHOLEBODY_N_KEYPOINTS = 133
