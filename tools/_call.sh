bash tools/final_evidence_r2.sh
