#!/usr/bin/env python
"""Generates tests/golden/mmc_reference.json: exit tallies of the REFERENCE's own MiniMC (NCrystal 4.4.2 built into
oracle/_ref by oracle/Makefile, driven through ncrystal_jsonquery ["mmc","run",...]) for the transport scenarios of
tests/_mmc.py, at 10x the statistics the tests run.  Needs /root/reference (through oracle/_ref) -- run in the
build container only; the resulting fixture travels to the GPU box.

    python tests/golden/make_golden_mmc.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _mmc import all_scenarios, reference_minimc, hists_from_json, NO_REFERENCE  # noqa: E402
from __graft_entry__ import CONFIGS  # noqa: E402

out = {}
for key, sc in all_scenarios().items():
    if key in NO_REFERENCE:
        continue
    n = 10 * sc.n
    js = reference_minimc(CONFIGS[sc.material], sc, nthreads=os.cpu_count() or 2, n=n)
    h, meta = hists_from_json(js, sc.tallies)
    out[key] = dict(n=n, cfgstr=CONFIGS[sc.material], geomcfg=sc.geomcfg, srccfg=sc.srccfg(n), enginecfg=sc.enginecfg(),
                    metadata=meta,
                    tallies={name: dict(content=v["total_content"].tolist(), errsq=v["total_errsq"].tolist(),
                                        class_integrals=v["content"].sum(axis=1).tolist(), stats=v["stats"])
                             for name, v in h.items()})
    print(key, meta)
json.dump(out, open(os.path.join(HERE, "mmc_reference.json"), "w"), indent=0)
