"""The Expr-side lowering binds to the drop-in: the reference's OWN ScalarFnCall::CreateSimplePredicates /
CreateOperate<T> (scalar-fn-call.cc:733-965) and And/OrPredicate::CreateSimplePredicates
(compound-predicates.cc:36-67), extracted at build time from /root/reference (never copied into the repo) and
compiled unmodified against stand-ins for the Impala classes they touch and against
host/exprs/simple-predicates.h; the resulting trees are lowered by the drop-in's Lower() and the postfix programs
checked.  See oracle/expr_binding_check.cc.  No GPU needed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "expr_binding_check")


def test_reference_expr_call_sites_compile_and_lower(lib_path):
    if os.path.exists("/root/reference/scalar-fn-call.cc"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "expr_binding"])
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/expr_binding_check not built (needs /root/reference at build time)")
    p = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "0 failure(s)" in p.stdout
