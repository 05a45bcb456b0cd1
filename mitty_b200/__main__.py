from mitty_b200.cli import cli

cli()
