"""`selections.txt` in the reference's format, so that a run of this package can be diffed against a run of the
reference with the reference's own tool.

The reference's training driver persists what each active-learning iteration selected as
`<experiment_dir>/run_%04d/selections.txt` (utils/saver.py:68-78): one line per selected image - the LMDB key decoded as
UTF-8 - followed, in region mode, by the image's regions as comma-separated `r,c,h,w` quadruples.  Its
`utils/compare_selections.py:4-25` walks the `run_*` folders two experiments have in common and prints how many NEW
lines of each iteration coincide.  `write_selections` produces that file from what the selector mirror returns;
`compare_selections` is the same comparison as a function (returns the counts instead of printing them).
"""
from __future__ import annotations

import os


def _text(path) -> str:
    return path.decode("utf-8") if isinstance(path, (bytes, bytearray)) else str(path)


def selection_lines(paths, regions=None):
    """Lines of selections.txt.  regions: None (image mode), a dict path -> [(r, c, h, w), ...] as
    create_region_maps returns it, or a list parallel to `paths`."""
    if regions is None:
        return [_text(p) + "\n" for p in paths]
    per_path = [regions[p] for p in paths] if isinstance(regions, dict) else list(regions)
    lines = []
    for p, region in zip(paths, per_path):
        flat = ",".join(",".join(str(int(v)) for v in r) for r in region)
        lines.append(_text(p) + "," + flat + "\n")
    return lines


def write_selections(experiment_dir: str, paths, regions=None, run: int | None = None) -> str:
    """Write `selections.txt` (into `experiment_dir/run_%04d` when `run` is given) and return its file name."""
    folder = experiment_dir if run is None else os.path.join(experiment_dir, "run_%04d" % run)
    os.makedirs(folder, exist_ok=True)
    filename = os.path.join(folder, "selections.txt")
    with open(filename, "w") as fptr:
        fptr.writelines(selection_lines(paths, regions))
    return filename


def compare_selections(folder_a: str, folder_b: str):
    """[(run folder, common new lines, new lines)] over the run folders both experiments have; raises ValueError when
    an iteration selected a different number of new lines in the two experiments (the reference's tool asserts)."""
    runs = sorted(set(d for d in os.listdir(folder_a) if os.path.isdir(os.path.join(folder_a, d)))
                  & set(d for d in os.listdir(folder_b) if os.path.isdir(os.path.join(folder_b, d))))
    seen_a, seen_b, out = set(), set(), []
    for run in runs:
        with open(os.path.join(folder_a, run, "selections.txt")) as f:
            new_a = set(f.readlines()) - seen_a
        with open(os.path.join(folder_b, run, "selections.txt")) as f:
            new_b = set(f.readlines()) - seen_b
        seen_a |= new_a
        seen_b |= new_b
        if len(new_a) != len(new_b):
            raise ValueError(f"unequal number of selections in {run}: {len(new_a)} vs {len(new_b)}")
        out.append((run, len(new_a & new_b), len(new_a)))
    return out
