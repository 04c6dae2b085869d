"""Stand-in for Biopython as far as scripts/python3/Context_genome_WG.py uses it (SeqIO.parse(handle, "fasta") records with
.description and .seq; Seq is imported but unused)."""
