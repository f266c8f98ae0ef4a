"""Test-only stand-in for the un-vendored timm==0.4.5 surface the reference imports.

Used ONLY by oracle/gen_golden.py and oracle/check_against_reference.py inside the build
container, so the reference's own modules (models/volo.py, loss/cross_entropy.py, prog/*)
can be imported from /root/reference and run on CPU.  Never imported by the product.
"""
