"""TEST INFRASTRUCTURE ONLY: CPU oracle for the nbnxm hot path. Never imported by gmxapi_b200."""
