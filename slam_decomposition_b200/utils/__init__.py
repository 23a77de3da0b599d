"""Helpers mirroring the in-scope parts of the reference's ``slam.utils`` package."""
