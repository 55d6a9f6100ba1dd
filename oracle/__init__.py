"""Test infrastructure (CPU oracle).  Never imported by the product package."""
