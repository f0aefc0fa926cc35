"""TEST INFRASTRUCTURE ONLY -- gate G0: run the reference's own 16 tests, unmodified, through
the gymnasium/pygame stand-ins of oracle/ref_loader.py (build container only)."""
import sys

import pytest

from oracle import ref_loader


def main() -> int:
    ref_loader.load_reference()
    sys.path.insert(0, ref_loader.REF_ROOT)
    return pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir=/tmp", "-o", "python_files=test_*.py",
                        ref_loader.REF_ROOT + "/tests"] + sys.argv[1:])


if __name__ == "__main__":
    raise SystemExit(main())
