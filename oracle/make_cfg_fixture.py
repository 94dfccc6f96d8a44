"""TEST INFRASTRUCTURE ONLY.  Executes the reference's own config files (plain Python,
read-only from /root/reference) and writes the keys the drop-ins consume to
``tests/golden/ref_cfg_keys.json``:

* cl_faster_rcnn_cfgs/incremental_task/cl_faster_rcnn_nsgp_repre_*.py:14-24  (task_id,
  train_task_split, offset, ignore_keys, previous_dir, ckpt_keywords, max_prototype, rr_thresh)
  and the ``roi_head`` keywords of its ``model`` dict (:65-70)
* cl_faster_rcnn_cfgs/_base_/schedules/schedule_1x_sgdnscl.py:19-23           (optim_wrapper)
* cl_faster_rcnn_cfgs/_base_/brnsrunetime.py:26                                (runner_type)

    python -m oracle.make_cfg_fixture          # needs /root/reference
"""
from __future__ import annotations

import json
import os

from .ref_loader import REFERENCE_ROOT

CFG_ROOT = os.path.join(REFERENCE_ROOT, "cl_faster_rcnn_cfgs")
TASK_FILES = ["cl_faster_rcnn_nsgp_repre_19_1_2.py", "cl_faster_rcnn_nsgp_repre_10_10_2.py",
              "cl_faster_rcnn_nsgp_repre_5_5_2.py", "cl_faster_rcnn_nsgp_repre_5_5_3.py",
              "cl_faster_rcnn_nsgp_repre_5_5_4.py"]
TOP_KEYS = ["task_id", "train_task_split", "offset", "ignore_keys", "previous_dir",
            "ckpt_keywords", "max_prototype", "rr_thresh"]
HEAD_KEYS = ["type", "previous_path", "task_id", "task_split", "max_prototype"]


def _exec(path):
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), ns)
    return ns


def collect():
    out = {"tasks": {}}
    for name in TASK_FILES:
        ns = _exec(os.path.join(CFG_ROOT, "incremental_task", name))
        ent = {k: ns[k] for k in TOP_KEYS if k in ns}
        head = ns["model"]["roi_head"]
        ent["roi_head"] = {k: head[k] for k in HEAD_KEYS if k in head}
        ent["bbox_head_type"] = head["bbox_head"]["type"]
        ent["_base_"] = ns["_base_"]
        out["tasks"][name] = ent
    sched = _exec(os.path.join(CFG_ROOT, "_base_", "schedules", "schedule_1x_sgdnscl.py"))
    out["optim_wrapper"] = sched["optim_wrapper"]
    run = _exec(os.path.join(CFG_ROOT, "_base_", "brnsrunetime.py"))
    out["runner_type"] = run["runner_type"]
    return out


def main():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(here, "tests", "golden", "ref_cfg_keys.json")
    with open(path, "w") as f:
        json.dump(collect(), f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
