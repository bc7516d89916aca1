"""Minimal stand-in for the absent `fvcore` package (TEST INFRASTRUCTURE ONLY).

The reference (/root/reference/setup.py:54) depends on fvcore, which is not installed in this
image and cannot be fetched.  Only three symbols touch the hot path; they are restated here from
their published behaviour.  Used solely by
oracle/gen_golden.py to import the reference in the authoring container.
"""
