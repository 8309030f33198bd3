"""`python -m ignnition_b200 main.py [args]`: run an unmodified IGNNITION user script.

Installs the `tensorflow` shim (only if TensorFlow is absent) and makes
`import framework_operations as ignnition` resolve to `ignnition_b200.framework_operations`, then
executes the script as `__main__` (the reference flow: `cd code && python3 main.py`)."""
import runpy
import sys

from . import framework_operations, tf_shim


def main():
    if len(sys.argv) < 2:
        print("usage: python -m ignnition_b200 main.py", file=sys.stderr)
        sys.exit(2)
    tf_shim.install()
    sys.modules.setdefault("framework_operations", framework_operations)
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
