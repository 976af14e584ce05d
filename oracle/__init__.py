"""ORACLE package: CPU restatement of the reference's denoising-step math. Test
infrastructure only -- never imported by sduss_b200/ (the product path)."""
