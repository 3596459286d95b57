"""Development-time helper: copies the PTM 65 nm BSIM4 model-card VALUES (public Predictive Technology Model data that
the reference ships as spice21opentechs/src/ptm65/{nmos,pmos}.yaml) into a JSON fixture, so that config C4
(BASELINE.json configs[3]) can be built on the GPU box where /root/reference does not exist.
Run from the repo root; needs /root/reference and PyYAML."""
import json
import yaml

out = {}
for name, mos_type in (("nmos", 0), ("pmos", 1)):
    card = yaml.safe_load(open(f"/root/reference/spice21opentechs/src/ptm65/{name}.yaml"))
    params = {}
    for k, v in card.items():
        if k in ("nmos", "pmos"):  # polarity flag; the Rust struct takes mos_type explicitly (SURVEY C4 note)
            continue
        params[k] = float(v)
    out[name] = {"mos_type": mos_type, "params": params}
json.dump(out, open("tests/golden/ptm65_cards.json", "w"), indent=0, sort_keys=True)
print({k: len(v["params"]) for k, v in out.items()})
