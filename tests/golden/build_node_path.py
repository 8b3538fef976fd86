"""Import helper: puts tests/dropin on sys.path so that tests can `import build_node`."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin"))
