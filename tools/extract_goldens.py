"""Extract the reference's own printed results for the hot path from its notebooks into
tests/golden/notebook_goldens.json.  Run in the build container (needs /root/reference; the GPU box never reads it).
Each golden records notebook, cell index, the cell source and the printed output verbatim."""
import json
import os
import sys

REF = "/root/reference/streamsculptor/examples"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "notebook_goldens.json")
# (id, notebook, a number string that identifies the output cell)
WANTED = [
    ("G1", "custom_potential.ipynb", "-0.186204177133592"),
    ("D1", "tests.ipynb", "-0.00238739"),
    ("D2", "tests.ipynb", "21.9793661"),
    ("D3_head", "tests.ipynb", "1.99947007e+01"),
    ("D3_tail", "tests.ipynb", "1.60397604e+01"),
    ("D4", "tests.ipynb", "-936.42809302"),
    ("D5", "StreamSubhaloExample.ipynb", "-7.23164146"),
    ("D7a", "linear_perturbation_stream.ipynb", "(1499, 50, 12)"),
    ("D8", "tests.ipynb", "9.96418614e-01"),
    ("B1", "tests.ipynb", "-1.10464897e-02"),                          # out_batch.ys[:, -1, 3]: 1000 orbits of integrate_orbit_batch_scan (cells 12, 15, 20, 22)
    ("R1", "tests.ipynb", "0.07882516"),                               # RestrictedNbody_generator.term at the two saved states of orbit 0 (cell 21)
    ("SS_impact", "StreamSubhaloExample.ipynb", "Impact location"),    # mean of a stream patch integrated back to the impact time (cells 1-7)
    ("SS_pot", "StreamSubhaloExample.ipynb", "Plummer subhalo potential"),   # SubhaloLinePotential[_Custom] values at (1,2,3), t = -850 (cell 9)
    ("OC", "OrphanChenab_mw_lmc_example.ipynb", "-2.19073895e+01"),   # stream_lead of the static-potential Orphan-Chenab stream (cells 3-6)
    ("S1", "tests.ipynb", "3.09774548e-04"),      # printed by an earlier revision (softened NFW): usable with the oracle's nfw(soft=1e-3)
]


def cell_outputs(cell):
    txt = []
    for o in cell.get("outputs", []):
        if "text" in o:
            txt.append("".join(o["text"]))
        elif "data" in o and "text/plain" in o["data"]:
            txt.append("".join(o["data"]["text/plain"]))
    return "\n".join(txt)


def main():
    res = []
    for gid, nbname, needle in WANTED:
        nb = json.load(open(os.path.join(REF, nbname)))
        hit = None
        for idx, cell in enumerate(nb["cells"]):
            if cell["cell_type"] != "code":
                continue
            out = cell_outputs(cell)
            if needle in out:
                hit = {"id": gid, "notebook": "streamsculptor/examples/" + nbname, "cell": idx, "source": "".join(cell["source"]), "output": out}
                break
        if hit is None:
            print("NOT FOUND:", gid, nbname, needle, file=sys.stderr)
            continue
        res.append(hit)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(res, open(OUT, "w"), indent=1)
    print(f"wrote {len(res)} goldens to {OUT}")


if __name__ == "__main__":
    main()
