module github.com/soypat/go-fem

go 1.18

require (
	github.com/soypat/lap v0.1.3
	gonum.org/v1/gonum v0.11.0
)
