"""CPU oracle for the ACF hot path -- TEST INFRASTRUCTURE, never imported by the product package."""
