"""Generate the golden fixtures under tests/golden/ from the reference tree (run in the build container).

    python tests/golden/make_golden.py

* ``metadata_fixture.json`` — the reference's own metadata fixture
  (``/root/reference/docs/_static/dummy_visualization_metadata.json``, used by
  ``test/utils/test_metadata.py:169-181``), copied byte for byte: it is test *data* (11 slice records of
  two observables), the input of the budget-replay known answers below and of the JSON round trip.
* ``metadata_expected.json`` — the expected ``accumulated_error`` / ``left_over_error_budget`` sequences
  the reference asserts to 10 places (``test/utils/test_metadata.py:183-265``), parsed out of that file.

/root/reference does not exist on the GPU box: tests read only the committed fixtures.
"""
import ast
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

shutil.copyfile(os.path.join(REF, "docs/_static/dummy_visualization_metadata.json"), os.path.join(HERE, "metadata_fixture.json"))

src = open(os.path.join(REF, "test/utils/test_metadata.py")).read()
tree = ast.parse(src)
lists: dict[str, list[float]] = {}
for node in ast.walk(tree):
    if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id.startswith("expected_obs_"):
        fn = next(
            f.name for f in ast.walk(tree) if isinstance(f, ast.FunctionDef) and f.lineno <= node.lineno <= f.end_lineno
        )
        lists[f"{fn}.{node.targets[0].id}"] = [float(ast.literal_eval(e)) for e in node.value.elts]
out = {
    "source": "test/utils/test_metadata.py:183-265",
    "accumulated_error": [lists["test_accumulated_error.expected_obs_1"], lists["test_accumulated_error.expected_obs_2"]],
    "left_over_error_budget": [lists["test_left_over_error_budget.expected_obs_1"], lists["test_left_over_error_budget.expected_obs_2"]],
}
json.dump(out, open(os.path.join(HERE, "metadata_expected.json"), "w"), indent=1)
print({k: len(v[0]) for k, v in out.items() if isinstance(v, list)})
