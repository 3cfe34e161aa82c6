"""
Simple demultiplexing with known genotypes -- the reference's examples/1-plain_demultiplexing.py on the B200 path,
without pysam:

    python examples/plain_demultiplexing.py /path/to/example_data

(example_data = test_bamfile.bam, test_barcodes.csv, test_genotypes.vcf as bundled with demuxalot)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from demuxalot_b200 import BarcodeHandler, Demultiplexer, ProbabilisticGenotypes, count_snps

data = Path(sys.argv[1] if len(sys.argv) > 1 else './example_data')

genotypes = ProbabilisticGenotypes(genotype_names=['Donor01', 'Donor02', 'Donor03', 'Donor04'])
genotypes.add_vcf(data / 'test_genotypes.vcf')
print(f'Loaded genotypes: {genotypes}')

barcode_handler = BarcodeHandler.from_file(data / 'test_barcodes.csv')
print(f'Loaded barcodes: {barcode_handler}')

snps = count_snps(
    bamfile_location=str(data / 'test_bamfile.bam'),
    chromosome2positions=genotypes.get_chromosome2positions(),
    barcode_handler=barcode_handler,
    joblib_n_jobs=1,
)
print('Collected SNPs: ')
for chromosome, calls in snps.items():
    print(f'Chromosome {chromosome}, {calls.n_snp_calls} calls in {calls.n_molecules} mols')

learnt_genotypes, posterior_probabilities = Demultiplexer.learn_genotypes(
    snps,
    genotypes=genotypes,
    barcode_handler=barcode_handler,
    doublet_prior=0.25,
)
print('Result:')
print(posterior_probabilities.round(3))
