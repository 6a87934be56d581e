"""The product package must not route through the oracle or any CPU fallback: nothing under pyradiomics_b200/ imports
or opens anything under oracle/ (only tests/, __graft_entry__.smoke()/build() and bench.py's CPU legs may)."""
import ast
import glob
import os

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
ORACLE_MODULES = {os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(ROOT, "oracle", "*.py"))}


def test_product_python_never_imports_the_oracle():
    assert {"cmatrices_oracle", "features_np", "pipeline", "shape_np", "build_ref", "ref_harness"} <= ORACLE_MODULES
    for path in glob.glob(os.path.join(ROOT, "pyradiomics_b200", "*.py")):
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name.split(".")[0] for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0 and node.module:
                names = [node.module.split(".")[0]]
            bad = [n for n in names if n in ORACLE_MODULES - {"pipeline"} or n == "oracle"]
            assert not bad, (path, bad)
            # `pipeline` is also the name of a product module (pyradiomics_b200/pipeline.py): only relative imports reach it
            assert "pipeline" not in names, (path, "absolute import of a module named like oracle/pipeline.py")
