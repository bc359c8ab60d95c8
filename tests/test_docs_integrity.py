"""CPU: the documents cite files (profiles, sources, tools); every cited path of this repository must exist,
so that the evidence a reader is pointed at is really there."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md")]


def cited_paths(text):
    for m in re.finditer(r"`([A-Za-z0-9_./{},*-]+)`", text):
        tok = m.group(1)
        if "*" in tok or "…" in tok or tok.startswith("/") or tok.endswith("/"):
            continue
        if not re.match(r"^(profiles|tools|tests|host|oracle|include|picnix_b200|csrc|experiments)/", tok) and \
                not re.match(r"^[a-z0-9_]+\.(cu|cuh|cpp|hpp|py|md|json|jsonl|txt|csv|sh)$", tok):
            continue
        # brace alternatives: r02_bench_n{2,4,8}.json
        alts = [tok]
        while any("{" in a for a in alts):
            nxt = []
            for a in alts:
                mm = re.search(r"\{([^{}]*)\}", a)
                if mm:
                    nxt += [a[:mm.start()] + o + a[mm.end():] for o in mm.group(1).split(",")]
                else:
                    nxt.append(a)
            alts = nxt
        yield from alts


def exists_somewhere(path):
    roots = ["", "profiles", "picnix_b200", os.path.join("picnix_b200", "csrc"),
             os.path.join("picnix_b200", "csrc", "experiments"), "host", os.path.join("host", "ref_binding"), "tools",
             "tests", "oracle", "include"]
    return any(os.path.exists(os.path.join(ROOT, r, path)) for r in roots)


@pytest.mark.parametrize("doc", DOCS)
def test_cited_files_exist(doc):
    text = open(os.path.join(ROOT, doc)).read()
    missing = sorted({p for p in cited_paths(text)
                      if not exists_somewhere(p)
                      # reference-tree files and build products are cited too; they are not part of the repository
                      and not re.match(r"^(nix|pic|example)/", p) and "_build" not in p and "_ref" not in p
                      and not p.endswith((".so", ".toml", ".msgpack")) and p not in ("history.txt", "config.toml", "main.cpp")})
    assert not missing, missing
