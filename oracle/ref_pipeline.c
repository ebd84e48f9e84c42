/* oracle/ref_pipeline.c -- TEST INFRASTRUCTURE ONLY (placeholder, filled in later). */
int oracle_ref_pipeline_version(void) { return 1; }
